// Micro-benchmark: per-SM throughput of the pipes the BA kernel leans on (B200, sm_100a).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, float* outf, int iters)
{
    double a = threadIdx.x * 1e-3 + 1.0, b = 1.0000001, c = 0.5, d = a + 1, e = a + 2, f = a + 3;
    float x = threadIdx.x * 1e-3f + 1.f, y = 1.0000001f, z = 0.5f, u = x + 1, v = x + 2, w = x + 3;
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { a = fma(a, b, c); d = fma(d, b, c); e = fma(e, b, c); f = fma(f, b, c); }          // DFMA
        if (OP == 1) { a = a + b; d = d + b; e = e + b; f = f + b; }                                       // DADD
        if (OP == 2) { a = a * b; d = d * b; e = e * b; f = f * b; }                                       // DMUL
        if (OP == 3) { x = fmaf(x, y, z); u = fmaf(u, y, z); v = fmaf(v, y, z); w = fmaf(w, y, z); }      // FFMA
        if (OP == 4) { a += (double)x; x += 1.f; d += (double)u; u += 1.f; e += (double)v; v += 1.f; f += (double)w; w += 1.f; } // F2F.F64.F32 + DADD + FADD
        if (OP == 5) { x += (float)a; a += 1.0; u += (float)d; d += 1.0; v += (float)e; e += 1.0; w += (float)f; f += 1.0; }   // F2F.F32.F64
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + d + e + f;
    outf[blockIdx.x * blockDim.x + threadIdx.x] = x + u + v + w;
}
template <int OP> void run(const char* name, int ops_per_iter)
{
    double* o; float* of;
    const int blocks = 148 * 4, threads = 512, iters = 4096;
    cudaMalloc(&o, blocks * threads * 8); cudaMalloc(&of, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(o, of, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(o, of, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * ops_per_iter;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-28s %8.3f ms  %8.2f Gop/s  ~%6.1f lanes/clk/SM (at %d MHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
    cudaFree(o); cudaFree(of);
}
int main()
{
    run<0>("DFMA", 4); run<1>("DADD", 4); run<2>("DMUL", 4); run<3>("FFMA", 4);
    run<4>("F2F.F64.F32 (+DADD,FADD)", 4); run<5>("F2F.F32.F64 (+FADD,DADD)", 4);
    return 0;
}
