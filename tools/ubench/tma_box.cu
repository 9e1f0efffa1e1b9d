// Finding: a box may start at any row / frame but its first byte must be 16-byte aligned (x % 4 == 0 for 32-bit words);
// the last variant below (x = 5) faults with 'illegal instruction'.
// Probe: which tensor-map shapes does cp.async.bulk.tensor accept for byte images seen as [frames][h][w*3/4] uint32?
// build: nvcc -std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a tools/ubench/tma_box.cu -o build/tma_box -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int z, int bytes, uint32_t* out)
{
    extern __shared__ __align__(128) uint8_t box[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&bar)), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_addr(box)),
                         "l"((uint64_t)&map), "r"(smem_addr(&bar)), "r"(x), "r"(y), "r"(z) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_addr(box)),
                         "l"((uint64_t)&map), "r"(smem_addr(&bar)), "r"(x), "r"(y) : "memory");
    }
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_addr(&bar)) : "memory");
        if (clock64() - t0 > 200000000ll) break;
    }
    if (threadIdx.x == 0) { out[0] = ok; out[1] = ((uint32_t*)box)[0]; out[2] = ((uint32_t*)box)[bytes / 4 - 1]; out[3] = smem_addr(box); }
}

int main()
{
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn)p;
    const int w = 512, h = 256, frames = 3;
    std::vector<uint32_t> host((size_t)w * 3 / 4 * h * frames);
    for (size_t i = 0; i < host.size(); i++) host[i] = (uint32_t)i;
    uint32_t *d, *out;
    cudaMalloc(&d, host.size() * 4);
    cudaMalloc(&out, 16);
    cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice);
    struct V { int rank, bw, bh, x, y, z, dim2; CUtensorMapL2promotion l2; } vs[] = {
        {2, 128, 40, 0, 0, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {2, 256, 64, 4, 7, 0, 0, CU_TENSOR_MAP_L2_PROMOTION_NONE},
        {3, 128, 40, 0, 0, 0, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {3, 128, 40, 8, 7, 1, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {3, 384, 48, 300, 230, 2, 3, CU_TENSOR_MAP_L2_PROMOTION_L2_128B}, {3, 256, 64, 12, 7, 1, 4096, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
        {3, 256, 40, 16, 7, 1, 4096, CU_TENSOR_MAP_L2_PROMOTION_NONE}, {3, 256, 40, 5, 7, 1, 4096, CU_TENSOR_MAP_L2_PROMOTION_NONE}};
    for (auto& v : vs) {
        CUtensorMap m;
        cuuint64_t dims[3] = {(cuuint64_t)w * 3 / 4, (cuuint64_t)h, (cuuint64_t)v.dim2};
        cuuint64_t strides[2] = {(cuuint64_t)w * 3, (cuuint64_t)w * h * 3};
        cuuint32_t box[3] = {(cuuint32_t)v.bw / 4, (cuuint32_t)v.bh, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, v.rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, v.l2,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int bytes = v.bw * v.bh;
        cudaMemset(out, 0, 16);
        if (v.rank == 3) { cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768); probe<3><<<1, 32, 32768>>>(m, v.x, v.y, v.z, bytes, out); }
        else { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768); probe<2><<<1, 32, 32768>>>(m, v.x, v.y, v.z, bytes, out); }
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t ho[4] = {0, 0, 0, 0};
        if (e == cudaSuccess) cudaMemcpy(ho, out, 16, cudaMemcpyDeviceToHost);
        const uint32_t want0 = (uint32_t)(((size_t)v.z * h + v.y) * (w * 3 / 4) + v.x);
        printf("rank %d box %dx%d at (%d,%d,%d) dim2 %d: encode %d, run %s, done %u, first %u (want %u), last %u, smem 0x%x\n", v.rank, v.bw, v.bh, v.x, v.y, v.z, v.dim2,
               (int)r, cudaGetErrorString(e), ho[0], ho[1], want0, ho[2], ho[3]);
        if (e != cudaSuccess) { printf("context lost, stopping\n"); return 1; }
    }
    return 0;
}
