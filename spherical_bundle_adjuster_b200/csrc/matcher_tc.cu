// matcher_tc.cu -- tensor-core matcher: tcgen05 (UMMA) bf16x3 distance contraction fed by TMA, fused with
// a per-row candidate reduction in the TMEM epilogue, then an exact fp32 re-rank.
//
// Replaces the knnMatch call of feature_matcher::match_two_image (feature_matcher.cpp:45) for large
// descriptor sets; the result is bit-identical to the exact SIMT kernel / cv::BFMatcher.
//
// Pipeline (all on the context's stream, no host synchronisation):
//   1. tc_prep      fp32 rows -> [hi | lo] bf16 planes (x = hi + lo + O(2^-17 |x|)) + squared norms.
//   2. tc_knn       persistent, warp-specialised, one CTA per SM (640 threads):
//                     warp 0   TMA producer: query block (A: 256 rows x {hi,lo}, 64 KB) once per query
//                              block; train tiles (B: 128 rows x {hi,lo}, 32 KB) through a 4-stage
//                              mbarrier ring.  One B tile feeds TWO M=128 MMA row-halves, which halves
//                              the L2->SM operand traffic per flop (at full tensor rate a 128-row query
//                              block would sit exactly on the L2 throughput cap).
//                     warp 1   MMA issuer: per tile 2 x 12 tcgen05.mma.kind::f16 (M128 N128 K16) =
//                              hi.hi + hi.lo + lo.hi accumulated in fp32 in TMEM (2 stages x 256 columns,
//                              so the epilogue of tile n overlaps the MMAs of tile n+1);
//                     warps 4-19 epilogue (16 warps = 4 TMEM lane quarters x 4 column slices of 64):
//                              tcgen05.ld 32 columns at a time, v = |b|^2 - 2 a.b (one FFMA), minimum per
//                              4-column chunk (3-input FMNMX), and a sorted top-4 of (chunk minimum, chunk
//                              id) per thread in registers; inserts are branch-free and guarded by warp
//                              votes.  The distance matrix never leaves the SM.
//                   Work = all (query block, train tile) pairs in query-major order, cut into one
//                   contiguous span per CTA; a span that ends inside a query block writes its partial
//                   top-4 lists to its own slots, so no inter-CTA merge or atomics are needed.
//   3. tc_rerank    per query row: merge the slots, evaluate the 16 rows of the 4 best chunks EXACTLY
//                   (OpenCV's fp32 order, shared with the SIMT kernel) -> top-2.  Every train row outside
//                   those chunks has approximate value >= B (the 4th best chunk minimum); with the proven
//                   bound |approx - exact| <= delta the row is final iff  d1^2 < B + |a|^2 - delta.
//                   Rows that fail the test (exact ties, >4 near-duplicates) are queued ...
//   4. tc_fallback  ... and re-scanned exactly over the whole train set, one CTA per queued row.
//
// Roofline: tensor pipe.  Algorithmic flops 2*D*N*M; the three bf16 products cost 3x that on the pipe.
#include <cuda.h>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "matcher_common.cuh"

namespace sba {

namespace tc {

constexpr int DIM = 64;            // descriptor length handled by this path
constexpr int KP = 2 * DIM;        // bf16 per prepared row: [hi(64) | lo(64)]
constexpr int BM = 256;            // query rows per block: two UMMA M=128 row-halves share every B tile
constexpr int BN = 128;            // train rows per tile (UMMA N)
constexpr int CHUNK = 8;           // columns per candidate chunk
constexpr int NCAND = 4;           // candidate chunks kept per list
constexpr int STAGES = 4;          // B smem ring
constexpr int EPI_WARPS = 16;      // 4 TMEM lane quarters x 4 column slices
constexpr int SUBSLOTS = 2;        // candidate lists per query row per CTA span (2 column slices per row-half)
constexpr int EPI_WARP0 = 3;      // first epilogue warp
constexpr int THREADS = (EPI_WARP0 + EPI_WARPS) * 32;   // warp0 TMA, warps 1-2 MMA issuers (even / odd tiles), warps 3-18 epilogue: 608 threads -> a budget of 104 registers each (95-96 used)
constexpr int A_KBLOCK_BYTES = BM * 128;   // one 64-wide 16-bit k-block of A: 32 KB
constexpr int B_KBLOCK_BYTES = BN * 128;   // 16 KB
// Two filter schemes, template parameter P = tensor products per result:
//   P = 3  bf16 split  x = hi + lo: hi.hi + hi.lo + lo.hi, error ~2^-16 |a||b|  (rows of 128 bf16: [hi | lo])
//   P = 1  plain fp16 rows: ONE product, error ~2^-10 |a||b|; a third of the tensor work, paid for with a wider safety
//          margin in the re-rank (more rows go to the exact fallback)
//   P = 2  128-d descriptors (feature_matcher.cpp:13, extended SURF): plain fp16 rows of 128 = two 64-wide k-blocks, one
//          product -- the same shared-memory footprint and the same epilogue as P = 3, 16 MMAs per tile
template <int P> struct Scheme {
    static constexpr int D = P == 2 ? 128 : 64;                            // descriptor length
    static constexpr int KBLOCKS = P == 1 ? 1 : 2;                         // 64-wide k-blocks per prepared row
    static constexpr int ROW_ELEMS = KBLOCKS * 64;                         // 16-bit elements per prepared row
    static constexpr int SMEM_A = KBLOCKS * A_KBLOCK_BYTES;
    static constexpr int B_STAGE_BYTES = KBLOCKS * B_KBLOCK_BYTES;
    static constexpr int SMEM_B = STAGES * B_STAGE_BYTES;
    static constexpr int SMEM_NB = EPI_WARPS * 64 * 4;                     // |b|^2 of each epilogue warp's columns
    static constexpr int SMEM_BYTES = SMEM_A + SMEM_B + SMEM_NB + 256 /*barriers: 20 x 8 B + the TMEM slot*/ + 1024 /*alignment slack*/;
};
constexpr int ACC_COLS = 2 * BN;   // TMEM columns per accumulator stage: row-half 0 | row-half 1
constexpr uint32_t TMEM_COLS = 512;
constexpr float DELTA_COEF = 4e-5f;  // bf16 x 3: |approx - exact| <= DELTA_COEF * (|a|^2 + max|b|^2): derivation at tc_rerank_kernel
constexpr float DELTA_COEF_FP16 = 1.05e-3f;  // fp16 x 1 (64-d and 128-d): same place
// Candidate keys (epilogue): a chunk minimum with the chunk's id in the low mantissa bits, so that ONE fp32 min / max
// moves value and id together.  9 bits: [8] = "old" flag, [7:3] tile inside the current 32-tile window, [2:0] chunk of
// the thread's 64 columns -- or, for entries that survived a window change, flag | list slot (the absolute chunk id of
// such an entry sits in a side register).  Truncating 9 mantissa bits moves a value by < 2^-14 of its magnitude.
constexpr uint32_t KEY_ID_MASK = 0x1FFu;
constexpr uint32_t KEY_OLD = 0x100u;
constexpr uint32_t KEY_CHUNK_MASK = 0x7u;      // the chunk bits alone (set first; window bits are added to the winners only)
constexpr int KEY_WINDOW = 32;                 // tiles per id window
constexpr uint32_t KEY_BIG = 0x7f7fffffu;      // FLT_MAX: an insert of this value is a no-op
constexpr uint32_t KEY_EMPTY = 0x7f000000u;    // value of an empty list entry (1.7e38, finite: keys must never be NaN)
constexpr float KEY_TRUNC_REL = 3.0f / 16384.0f;   // bound test margin for the truncation + id bits, relative to |B| (see tc_rerank_kernel)
constexpr float PAD_NORM = 1e30f;              // |b|^2 of padding train rows: huge but finite (inf would turn keys into NaN)

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol bug traps (CUDA error) after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t ok = 0;
    long long t0 = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) break;
        // the clock is only read once a wait has failed (the common case returns on the first try)
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 60000000000ll) __trap();   // ~30 s: a protocol bug, not ordinary skew
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// ---- CTA-pair (cta_group::2) variants: two CTAs of one cluster (one TPC) run every MMA together -- M = 256, each CTA's tensor
// core works on its own 128 rows of A, the B operand is split in halves of 64 columns, one half in each CTA's shared memory,
// and read ONCE for both tensor cores.  Barriers that gate the MMA issue live in the even CTA (the leader, the only one that
// issues); an address of the executing CTA with the peer bit cleared names the leader's copy of the same object.
constexpr uint32_t PEER_BIT = 0x01000000u;
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx_leader(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(smem_u32(bar) & ~PEER_BIT), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & ~PEER_BIT) : "memory");
}
__device__ __forceinline__ void mbar_arrive_peer(uint64_t* bar)   // called by the leader: the odd CTA's copy
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) | PEER_BIT) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"((uint64_t)map), "r"(smem_u32(bar) & ~PEER_BIT), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar)   // arrives on the barrier at this offset in BOTH CTAs
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, M128 x N128 x K16.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// The same issued for a CTA pair: M256 (128 rows in each CTA) x N128 (64 columns from each CTA's shared memory) x K16.
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major, 128-byte-swizzled operand whose rows are 128 B apart and whose 8-row groups are 1024 B apart
// (exactly what a TMA box {64 bf16, rows} with CU_TENSOR_MAP_SWIZZLE_128B writes).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N=128, M=128.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// the same with fp16 operands (A / B format fields 0)
constexpr uint32_t IDESC_F16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
// CTA pair: M = 256
constexpr uint32_t IDESC_PAIR = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t IDESC_F16_PAIR = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

#define TMEM_LD_X32(r, taddr)                                                                                                  \
    asm volatile(                                                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                              \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),   \
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),     \
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),     \
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                                    \
        : "r"(taddr)                                                                                                           \
        : "memory")

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- work partition (shared by tc_knn and tc_rerank) ------------------------------------------------
struct Partition {
    int nqb, ntb, n_ctas;
    long long T;
    // Entering a query block costs a CTA about `bcost` tiles' worth of time (the A operand reload drains the MMA
    // pipeline and the candidate lists start empty again), so spans are cut in COST space: every block start
    // weighs bcost tiles.  Position x in cost space -> tile index; the weight sits before the block's first tile.
    int bcost;
    __host__ __device__ long long cost_total() const { return T + (long long)bcost * nqb; }
    __host__ __device__ long long tile_at(long long x) const
    {
        const long long per = ntb + bcost, q = x / per, rem = x - q * per;
        return q * ntb + (rem > bcost ? rem - bcost : 0);
    }
    __host__ __device__ long long start(int c) const { return tile_at((long long)c * cost_total() / n_ctas); }
    // the CTA whose span contains tile t
    __host__ __device__ int cta_of(long long t) const
    {
        const long long x = t + (long long)bcost * (t / ntb + 1);
        int k = (int)(x * n_ctas / cost_total());
        if (k >= n_ctas) k = n_ctas - 1;
        while (k + 1 < n_ctas && start(k + 1) <= t) k++;
        while (k > 0 && start(k) > t) k--;
        return k;
    }
    __host__ __device__ int max_slots() const { return (n_ctas + nqb - 1) / nqb + 2; }
};

// ---- 1. prepare ------------------------------------------------------------------------------------------
// D/4 threads per row, 4 floats each.  Rows >= n are padding: zeros, norm = PAD_NORM (train) so they never win.
template <int D>
__device__ __forceinline__ float prep_row(const float* __restrict__ x, int n, int row, int part, __nv_bfloat16* __restrict__ out, float* __restrict__ norm,
                                          float pad_norm, __half* __restrict__ out16)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row < n) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)row * D) + part);
    const float f[4] = {v.x, v.y, v.z, v.w};
    __align__(8) __nv_bfloat16 hi[4];
    __align__(8) __nv_bfloat16 lo[4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        hi[k] = __float2bfloat16_rn(f[k]);
        lo[k] = __float2bfloat16_rn(f[k] - __bfloat162float(hi[k]));
        s = fmaf(f[k], f[k], s);
    }
#pragma unroll
    for (int o = D / 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, D / 4);
    if (out) {   // D == 64 only
        __nv_bfloat16* dst = out + (size_t)row * KP;
        *reinterpret_cast<uint2*>(dst + part * 4) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(dst + DIM + part * 4) = *reinterpret_cast<const uint2*>(lo);
    }
    if (out16) {   // plain fp16 rows (the fp16 filter): see the range note at tc_rerank_kernel
        __align__(8) __half h[4];
#pragma unroll
        for (int k = 0; k < 4; k++) h[k] = __float2half_rn(f[k]);
        *reinterpret_cast<uint2*>(out16 + (size_t)row * D + part * 4) = *reinterpret_cast<const uint2*>(h);
    }
    if (part == 0) norm[row] = row < n ? s : pad_norm;
    return row < n ? s : 0.f;   // for the caller's running maximum (padding rows do not count)
}

#ifdef SBA_TC_TRACE
// Debug build only: [0] = last thread of tc_prep_kernel to finish, [1] = first thread of tc_rerank_kernel to start (globaltimer, ns)
__device__ unsigned long long g_tc_edge[2];
__device__ __forceinline__ unsigned long long gtime_early()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// Both descriptor sets in one launch: rows [0, nq_pad) are queries, the rest train rows.
// out*: bf16 hi|lo rows (P = 3) or NULL; out16*: fp16 rows (P = 1) or NULL.  na_max / nb_max: largest squared norms.
template <int D>
__global__ void tc_prep_kernel(const float* __restrict__ q, int nq, int nq_pad, __nv_bfloat16* __restrict__ outA, float* __restrict__ na,
                               const float* __restrict__ t, int nt, int nt_pad, __nv_bfloat16* __restrict__ outB, float* __restrict__ nb,
                               float* __restrict__ nb_max /* running maximum of the finite train norms */, float* __restrict__ na_max,
                               __half* __restrict__ out16A, __half* __restrict__ out16B)
{
    pdl_trigger();   // the distance kernel may be scheduled while this one drains (it waits before touching memory)
    __shared__ float s_max[2][8];
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int row = gid / (D / 4), part = gid % (D / 4);   // D/4 threads per row, 4 floats each; a row never straddles two warps or the two sets
    float mq = 0.f, mt = 0.f;   // squared norms are non-negative: 0 is the neutral element
    if (row < nq_pad) mq = prep_row<D>(q, nq, row, part, outA, na, 0.f, out16A);
    else if (row < nq_pad + nt_pad) mt = prep_row<D>(t, nt, row - nq_pad, part, outB, nb, PAD_NORM, out16B);
    // running maxima of the norms: one atomic per CTA and set instead of one per row (32 768 same-address atomics serialise in L2)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mq = fmaxf(mq, __shfl_xor_sync(0xffffffffu, mq, o));
        mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, o));
    }
    if ((threadIdx.x & 31) == 0) { s_max[0][threadIdx.x >> 5] = mq; s_max[1][threadIdx.x >> 5] = mt; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float m = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) m = fmaxf(m, s_max[threadIdx.x][w]);
        float* dst = threadIdx.x == 0 ? na_max : nb_max;
        if (dst && m > 0.f) atomicMax((int*)dst, __float_as_int(m));   // non-negative floats order like ints
    }
#ifdef SBA_TC_TRACE
    if (threadIdx.x == 0) atomicMax(&g_tc_edge[0], gtime_early());
#endif
}

// ---- 2. the tensor-core kernel ---------------------------------------------------------------------------
struct Cand {
    float v[NCAND];
    int id[NCAND];
};

__device__ __forceinline__ void cand_insert(Cand& c, float m, int id)
{
    if (m < c.v[3]) {
        if (m < c.v[2]) {
            c.v[3] = c.v[2]; c.id[3] = c.id[2];
            if (m < c.v[1]) {
                c.v[2] = c.v[1]; c.id[2] = c.id[1];
                if (m < c.v[0]) { c.v[1] = c.v[0]; c.id[1] = c.id[0]; c.v[0] = m; c.id[0] = id; }
                else { c.v[1] = m; c.id[1] = id; }
            } else { c.v[2] = m; c.id[2] = id; }
        } else { c.v[3] = m; c.id[3] = id; }
    }
}

// Branch-free variant for the epilogue (same result): place at the tail if it beats the tail, then
// bubble up with three compare-exchanges.  Strict '<' keeps the earlier chunk ahead on ties.
__device__ __forceinline__ void cand_insert_bf(Cand& c, float m, int id)
{
    const bool p = m < c.v[3];
    c.v[3] = p ? m : c.v[3];
    c.id[3] = p ? id : c.id[3];
#pragma unroll
    for (int k = 3; k > 0; k--) {
        const bool sw = c.v[k] < c.v[k - 1];
        const float tv = c.v[k - 1];
        const int ti = c.id[k - 1];
        c.v[k - 1] = sw ? c.v[k] : tv;
        c.id[k - 1] = sw ? c.id[k] : ti;
        c.v[k] = sw ? tv : c.v[k];
        c.id[k] = sw ? ti : c.id[k];
    }
}

// (d0, d1) = -2 * (a0, a1) + (c0, c1) with one fma.rn.f32x2
__device__ __forceinline__ void ffma2(float& d0, float& d1, uint32_t a0, uint32_t a1, float c0, float c1)
{
    uint64_t a, c, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(0xC0000000C0000000ull), "l"(c));   // -2.0f twice
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

__device__ __forceinline__ void cand_reset(Cand& c)
{
#pragma unroll
    for (int k = 0; k < NCAND; k++) { c.v[k] = __int_as_float(0x7f800000); c.id[k] = -1; }
}

// ---- key lists: the four smallest keys of everything inserted so far, sorted, in four registers --------------
// Insert = a 7-instruction min/max chain, no predicates, no moves.  Keys are unique (distinct ids), so nothing is lost.
__device__ __forceinline__ void key_insert(float (&k)[NCAND], float x)
{
    const float t = fmaxf(k[0], x); k[0] = fminf(k[0], x);
    const float u = fmaxf(k[1], t); k[1] = fminf(k[1], t);
    const float w = fmaxf(k[2], u); k[2] = fminf(k[2], u);
    k[3] = fminf(k[3], w);
}

__device__ __forceinline__ void key_reset(float (&k)[NCAND], int (&abs_id)[NCAND])
{
#pragma unroll
    for (int i = 0; i < NCAND; i++) { k[i] = __uint_as_float(KEY_EMPTY | KEY_OLD | (uint32_t)i); abs_id[i] = -1; }
}

// End of an id window: every entry becomes "old" -- its absolute chunk id moves to the side register of its slot.
// chunk_base = absolute id of chunk 0 of the window's first tile for this thread's column half.
__device__ __forceinline__ void key_flush(float (&k)[NCAND], int (&abs_id)[NCAND], int chunk_base)
{
    int na[NCAND];
#pragma unroll
    for (int i = 0; i < NCAND; i++) {
        const uint32_t kb = __float_as_uint(k[i]);
        const uint32_t idf = kb & KEY_ID_MASK;
        const uint32_t sl = idf & 3u;
        int old = abs_id[0];
        old = sl == 1u ? abs_id[1] : old;
        old = sl == 2u ? abs_id[2] : old;
        old = sl == 3u ? abs_id[3] : old;
        const int fresh = chunk_base + (int)((idf >> 3) & 31u) * (BN / CHUNK) + (int)(idf & 7u);
        na[i] = (idf & KEY_OLD) ? old : fresh;
        k[i] = __uint_as_float((kb & ~KEY_ID_MASK) | KEY_OLD | (uint32_t)i);
    }
#pragma unroll
    for (int i = 0; i < NCAND; i++) abs_id[i] = na[i];
}

#ifdef SBA_TC_TRACE
// Debug build only (make TRACE=1): per-CTA time stamps, read back by tools/tc_trace.py through sba_tc_trace_buffer().
__device__ unsigned long long g_tc_trace[148 * 16];
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define TC_TRACE(slot) do { g_tc_trace[(blockIdx.x % 148) * 16 + (slot)] = gtime(); } while (0)
#else
#define TC_TRACE(slot) do { } while (0)
#endif

// PAIR: the CTAs 2p and 2p+1 form a cluster and walk ONE span of (query super-block, train tile) together: CTA rank r owns query
// block 2 * super-block + r (its own A operand, accumulators, epilogue and candidate lists, exactly as in the single-CTA kernel)
// and loads rows 64r .. 64r+63 of every train tile; the leader (rank 0) issues cta_group::2 MMAs for both.  The partition,
// the list slots and the re-rank's span table are then in units of pairs and super-blocks (`part` arrives in those units).
template <int P, bool PAIR = false>
__global__ void __launch_bounds__(THREADS, 1)
tc_knn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const float* __restrict__ nb,
              Partition part, float4* __restrict__ cand_v, int4* __restrict__ cand_id, int slots)
{
    using S = Scheme<P>;
    constexpr int SMEM_A = S::SMEM_A, SMEM_B = S::SMEM_B, SMEM_NB = S::SMEM_NB, B_STAGE_BYTES = S::B_STAGE_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment for the 128B-swizzle atoms; offset arithmetic keeps the shared address space
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                       // [kblocks][256 rows][128 B]
    uint8_t* sB = smem + SMEM_A;              // [STAGES][kblocks][128 rows][128 B]
    float* sNB = (float*)(smem + SMEM_A + SMEM_B);  // [EPI_WARPS][64 (x2)]: |b|^2 (and row scales) of each epilogue warp's columns
    uint64_t* bars = (uint64_t*)(smem + SMEM_A + SMEM_B + SMEM_NB);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* b_full = bars + 2;               // [STAGES]
    uint64_t* b_empty = bars + 2 + STAGES;     // [STAGES]
    // accumulator barriers per (stage, row-half): the epilogue warps of row-half 0 start on a tile while the MMAs of
    // row-half 1 are still running, which also keeps the two halves' warps out of phase on every scheduler
    uint64_t* acc_full = bars + 2 + 2 * STAGES;   // [2 stages][2 halves]
    uint64_t* acc_empty = bars + 6 + 2 * STAGES;  // [2 stages][2 halves]
    uint64_t* turn = bars + 10 + 2 * STAGES;      // [2]: issue-order token between the two MMA issuer warps
    uint32_t* tmem_slot = (uint32_t*)(bars + 12 + 2 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    if (threadIdx.x == 0) TC_TRACE(0);   // CTA start
    const int rank = PAIR ? (int)cluster_ctarank() : 0;             // CTA inside its pair
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // what the partition counts: pairs or CTAs
    const int qmul = PAIR ? 2 : 1;                                   // query blocks per partition block
    // this CTA's span of tiles, walked with 32-bit incremental (query block, train tile) indices
    const int t_begin = (int)part.start(unit), n_tiles = (int)part.start(unit + 1) - t_begin;
    const int ntb = part.ntb;
    const int qb0 = t_begin / ntb, tb0 = t_begin - qb0 * ntb;

    if (threadIdx.x == 0) {
        // pair mode: the leader's "full" barriers collect one arrival (+ its bytes) from each CTA's producer, its accumulator
        // "empty" barriers one from every epilogue warp of both CTAs; "empty" / "ready" signals go out to both CTAs at once
        mbar_init(a_full, PAIR ? 2 : 1);
        mbar_init(a_empty, 2);   // one arrival per MMA issuer warp
        for (int s = 0; s < STAGES; s++) { mbar_init(b_full + s, PAIR ? 2 : 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 4; s++) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, PAIR ? EPI_WARPS : EPI_WARPS / 2); }
        mbar_init(turn + 0, 1);
        mbar_init(turn + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // the peer's barriers exist before anything arrives on them
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();   // barriers and TMEM are set up; the operands (tc_prep_kernel's output) are touched only from here on
    if (threadIdx.x == 0) TC_TRACE(1);   // set-up done

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int qb = qb0, tb = tb0, seg = 0, s = 0;
            uint32_t ring_phase = 0;
            bool new_seg = true;
            for (int n = 0; n < n_tiles; n++) {
                if (new_seg) {
                    mbar_wait(a_empty, (seg & 1) ^ 1);
                    const int arow = (qb * qmul + rank) * BM;   // this CTA's own query block
                    if (PAIR) {
                        mbar_expect_tx_leader(a_full, SMEM_A);
                        tma_load_2d_pair(sA, &map_a, a_full, 0, arow);
                        if (P != 1) tma_load_2d_pair(sA + A_KBLOCK_BYTES, &map_a, a_full, 64, arow);
                    } else {
                        mbar_expect_tx(a_full, SMEM_A);
                        tma_load_2d(sA, &map_a, a_full, 0, arow);
                        if (P != 1) tma_load_2d(sA + A_KBLOCK_BYTES, &map_a, a_full, 64, arow);
                    }
                    seg++;
                }
                mbar_wait(b_empty + s, ring_phase ^ 1);
#ifdef SBA_TC_EXP_NOTMA   // experiment (trace builds): B tiles are loaded for the first ring pass only -- WRONG results, timing probe
                if (n >= STAGES) { mbar_arrive(b_full + s); } else
#endif
                {
                uint8_t* dst = sB + s * B_STAGE_BYTES;
                if (PAIR) {   // this CTA's half of the tile (map_b is the 64-row box here): rows 0..63 of each k-block's region
                    mbar_expect_tx_leader(b_full + s, B_STAGE_BYTES / 2);
                    tma_load_2d_pair(dst, &map_b, b_full + s, 0, tb * BN + rank * (BN / 2));
                    if (P != 1) tma_load_2d_pair(dst + B_KBLOCK_BYTES, &map_b, b_full + s, 64, tb * BN + rank * (BN / 2));
                } else {
                    mbar_expect_tx(b_full + s, B_STAGE_BYTES);
                    tma_load_2d(dst, &map_b, b_full + s, 0, tb * BN);
                    if (P != 1) tma_load_2d(dst + B_KBLOCK_BYTES, &map_b, b_full + s, 64, tb * BN);
                }
                }
                if (++s == STAGES) { s = 0; ring_phase ^= 1; }
                new_seg = (++tb == ntb);
                if (new_seg) { tb = 0; qb++; }
            }
        }
        __syncwarp();
    } else if ((warp == 1 || warp == 2) && rank == 0) {
        // ===== MMA issuers (pair mode: in the leader CTA only): warp 1 takes the even tiles of the span (accumulator stage 0), warp 2 the odd ones =====
        // The tensor pipe's queue is shallow: measured with the probe builds (Makefile `trace EXP=...`), a single
        // issuing warp costs 0.27 us of idle pipe per tile -- its barrier waits, fences and commits -- on top of
        // 35 ns per MMA.  Two issuers hide each other's per-tile overhead; tiles are independent (own accumulator
        // stage, own B stage), so no ordering between the two warps is needed.
        // Each warp walks the tile loop in lock-step (waits included) so every descriptor is warp-uniform and
        // lives in uniform registers; only the tcgen05 instructions themselves are issued by one elected lane.
        constexpr uint32_t ID_BF = PAIR ? IDESC_PAIR : IDESC, ID_F16 = PAIR ? IDESC_F16_PAIR : IDESC_F16;
        auto UMMA = [](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
            if (PAIR) umma_bf16_pair(d, da, db, idesc, accumulate);
            else umma_bf16(d, da, db, idesc, accumulate);
        };
        auto COMMIT = [](uint64_t* bar) {
            if (PAIR) tcgen05_commit_pair(bar);
            else tcgen05_commit(bar);
        };
        uint32_t is_leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(is_leader));
        const int my_parity = warp - 1;
        int tb = tb0, seg = 0, s = 0;
        uint32_t ring_phase = 0;
        bool new_seg = true, issued_in_block = false;
        const uint32_t sa = smem_u32(sA);
        // row-half h of the query block: rows 128h.. of each A k-block
        const uint64_t dA_hi0 = make_smem_desc(sa), dA_hi1 = make_smem_desc(sa + 128 * 128);
        const uint64_t dA_lo0 = make_smem_desc(sa + A_KBLOCK_BYTES), dA_lo1 = make_smem_desc(sa + A_KBLOCK_BYTES + 128 * 128);
        const uint64_t dB0 = make_smem_desc(smem_u32(sB));
        for (int n = 0; n < n_tiles; n++) {
            const bool mine = (n & 1) == my_parity;
            if (new_seg) seg++;
            if (mine) {
                if (!issued_in_block) mbar_wait(a_full, (seg - 1) & 1);   // first tile of this warp in the query block
                const int acc = n & 1;
                mbar_wait(acc_empty + 2 * acc, (uint32_t)(((n >> 1) & 1) ^ 1));
                mbar_wait(b_full + s, ring_phase);
                // issue order: the tensor pipe runs MMAs in the order they arrive, so tile n+1 must not be issued before ALL of
                // tile n has been (interleaved tiles would finish together and starve the epilogue's two-stage hand-over).
                // The other warp passes the token as soon as its last MMA is queued; everything above already happened.
                mbar_wait(turn + my_parity, (uint32_t)((n >> 1) & 1) ^ (uint32_t)(my_parity ^ 1));
                tcgen05_fence_after();
                if (n == 0 && is_leader) TC_TRACE(2);              // first operands landed
                if (n == n_tiles - 1 && is_leader) TC_TRACE(3);    // last tile's MMAs about to issue
                // descriptors address 16-byte units: stage stride and k-block stride are plain adds
                const uint64_t dB_hi = dB0 + (uint64_t)(s * (B_STAGE_BYTES >> 4));
                const uint64_t dB_lo = dB_hi + (B_KBLOCK_BYTES >> 4);
                const uint32_t d0 = tmem_base + (uint32_t)(acc * ACC_COLS), d1 = d0 + BN;
                if (is_leader) {
                    // hi.hi + hi.lo + lo.hi ; each 64-wide k-block is four K=16 steps, 32 B apart
#pragma unroll
                    for (int k = 0; k < 4; k++) UMMA(d0, dA_hi0 + 2 * k, dB_hi + 2 * k, P == 3 ? ID_BF : ID_F16, k > 0);
#ifndef SBA_TC_EXP_1PROD   // experiment (trace builds): one product only -- WRONG results, timing probe
                    if (P == 3) {
#pragma unroll
                        for (int k = 0; k < 4; k++) UMMA(d0, dA_hi0 + 2 * k, dB_lo + 2 * k, ID_BF, 1);
#pragma unroll
                        for (int k = 0; k < 4; k++) UMMA(d0, dA_lo0 + 2 * k, dB_hi + 2 * k, ID_BF, 1);
                    } else if (P == 2) {   // second half of the 128-long rows
#pragma unroll
                        for (int k = 0; k < 4; k++) UMMA(d0, dA_lo0 + 2 * k, dB_lo + 2 * k, ID_F16, 1);
                    }
#endif
                    COMMIT(acc_full + 2 * acc);      // row-half 0 ready for its epilogue warps
                }
                mbar_wait(acc_empty + 2 * acc + 1, (uint32_t)(((n >> 1) & 1) ^ 1));
                tcgen05_fence_after();
                if (is_leader) {
#pragma unroll
                    for (int k = 0; k < 4; k++) UMMA(d1, dA_hi1 + 2 * k, dB_hi + 2 * k, P == 3 ? ID_BF : ID_F16, k > 0);
#ifndef SBA_TC_EXP_1PROD
                    if (P == 3) {
#pragma unroll
                        for (int k = 0; k < 4; k++) UMMA(d1, dA_hi1 + 2 * k, dB_lo + 2 * k, ID_BF, 1);
#pragma unroll
                        for (int k = 0; k < 4; k++) UMMA(d1, dA_lo1 + 2 * k, dB_hi + 2 * k, ID_BF, 1);
                    } else if (P == 2) {
#pragma unroll
                        for (int k = 0; k < 4; k++) UMMA(d1, dA_lo1 + 2 * k, dB_lo + 2 * k, ID_F16, 1);
                    }
#endif
                    mbar_arrive(turn + (my_parity ^ 1));     // the other issuer may queue the next tile
                    COMMIT(b_empty + s);             // B stage free once these MMAs have read it
                    COMMIT(acc_full + 2 * acc + 1);  // row-half 1 ready
                }
                issued_in_block = true;
            }
            if (++s == STAGES) { s = 0; ring_phase ^= 1; }
            new_seg = (++tb == ntb);
            if (new_seg) tb = 0;
            if (new_seg || n + 1 == n_tiles) {   // last tile of this query block in the span: A may be overwritten once BOTH
                                                 // warps' MMAs on it are complete (a_empty counts two arrivals)
                if (is_leader) {
                    if (issued_in_block) COMMIT(a_empty);
                    else {                       // this warp had no tile in the block
                        mbar_arrive(a_empty);
                        if (PAIR) mbar_arrive_peer(a_empty);
                    }
                }
                issued_in_block = false;
            }
            __syncwarp();
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue warps: TMEM lane quarter = warp % 4 (hardware rule), column slice = (warp - EPI_WARP0) / 4 =====
        const int quarter = warp & 3;
        const int slice = (warp - EPI_WARP0) >> 2;    // 0..3: 64 accumulator columns each
        const int half = slice >> 1;                  // which M=128 row-half those columns belong to
        const int csub = slice & 1;                   // which 64 train columns of the tile
        const int row = half * 128 + quarter * 32 + lane;   // row inside the 256-row query block
        float* wnb = sNB + (warp - EPI_WARP0) * 64;   // this warp's |b|^2 staging (warp-synchronous)
        const float* nb_col = nb + csub * 64 + lane;
        float key[NCAND];
        int abs_id[NCAND];
        key_reset(key, abs_id);
        // the chunk numbers 1..7 in registers the compiler cannot fold: (m & ~7) | c is then ONE LOP3 (immediate mask, register id)
        uint32_t creg[8];
#pragma unroll
        for (int c = 0; c < 8; c++) creg[c] = (uint32_t)c + (blockDim.y - 1u);   // blockDim.y == 1: a zero the compiler cannot see
        int qb = qb0, tb = tb0;
        int tw = 0;                                   // tile inside the current id window
        int chunk_base = tb0 * (BN / CHUNK) + csub * 8;   // absolute id of the window's first chunk for this column half
        float nb0 = __ldg(nb_col + (size_t)tb * BN), nb1 = __ldg(nb_col + (size_t)tb * BN + 32);
        for (int n = 0; n < n_tiles; n++) {
            const int acc = n & 1;
            __syncwarp();                             // every lane is done reading the previous tile's values
            wnb[lane] = nb0;
            wnb[lane + 32] = nb1;
            __syncwarp();
            const int tb_next = (tb + 1 == ntb) ? 0 : tb + 1;
            if (n + 1 < n_tiles) {                    // prefetch the next tile's norms behind this tile's math
                nb0 = __ldg(nb_col + (size_t)tb_next * BN);
                nb1 = __ldg(nb_col + (size_t)tb_next * BN + 32);
            }
            mbar_wait(acc_full + 2 * acc + half, (uint32_t)((n >> 1) & 1));
            tcgen05_fence_after();
            if (warp == EPI_WARP0 && lane == 0 && n == 0) TC_TRACE(4);             // first accumulator ready
            if (warp == EPI_WARP0 && lane == 0 && n == n_tiles - 1) TC_TRACE(5);   // last accumulator ready
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * ACC_COLS + slice * 64);
            uint32_t buf[64];
#ifdef SBA_TC_EXP_NOLDTM   // experiment (trace builds): the accumulators are never read -- WRONG results, timing probe
#pragma unroll
            for (int i = 0; i < 64; i++) buf[i] = taddr + i;
#else
            TMEM_LD_X32(buf, taddr);
            TMEM_LD_X32((buf + 32), taddr + 32);
            tmem_ld_wait();
#endif
            // release the accumulator stage as soon as its values sit in registers
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_leader(acc_empty + 2 * acc + half);
                else mbar_arrive(acc_empty + 2 * acc + half);
            }
#ifdef SBA_TC_EXP_NOEPI   // experiment (trace builds): accumulators are read and dropped -- WRONG results, timing probe
            if (buf[0] == 0x12345678u && buf[63] == 0x9abcdef0u) key[0] = 0.f;
#else
            // eight chunk keys: minimum of 8 columns of |b|^2 - 2 a.b, chunk number in the three lowest mantissa bits
            float ck[8];
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const float4 na4 = *reinterpret_cast<const float4*>(wnb + c * 8);
                const float4 nb4 = *reinterpret_cast<const float4*>(wnb + c * 8 + 4);
                float v0, v1, v2, v3, v4, v5, v6, v7;   // packed fp32 FMAs, same rounding as eight fmaf
                ffma2(v0, v1, buf[c * 8 + 0], buf[c * 8 + 1], na4.x, na4.y);
                ffma2(v2, v3, buf[c * 8 + 2], buf[c * 8 + 3], na4.z, na4.w);
                ffma2(v4, v5, buf[c * 8 + 4], buf[c * 8 + 5], nb4.x, nb4.y);
                ffma2(v6, v7, buf[c * 8 + 6], buf[c * 8 + 7], nb4.z, nb4.w);
                const float m = fminf(fminf(fminf(fminf(v0, v1), v2), fminf(fminf(v3, v4), v5)), fminf(v6, v7));
                ck[c] = __uint_as_float((__float_as_uint(m) & ~KEY_CHUNK_MASK) | creg[c]);
            }
            // smallest (gk) and second smallest (m2) of the eight chunk keys: 18 min/max
            const float lo0 = fminf(ck[0], ck[1]), hi0 = fmaxf(ck[0], ck[1]);
            const float lo1 = fminf(ck[2], ck[3]), hi1 = fmaxf(ck[2], ck[3]);
            const float lo2 = fminf(ck[4], ck[5]), hi2 = fmaxf(ck[4], ck[5]);
            const float lo3 = fminf(ck[6], ck[7]), hi3 = fmaxf(ck[6], ck[7]);
            const float l01 = fminf(lo0, lo1), h01 = fmaxf(lo0, lo1);
            const float l23 = fminf(lo2, lo3), h23 = fmaxf(lo2, lo3);
            const float gk = fminf(l01, l23);
            // window bits go onto the winners only: key = value bits above the id field | tile-in-window | chunk
            const uint32_t idb = (uint32_t)tw << 3;                        // warp-uniform
            const uint32_t keep = ~KEY_ID_MASK | KEY_CHUNK_MASK;
            const float gkf = __uint_as_float((__float_as_uint(gk) & keep) | idb);
            // Common case: at most the tile's best chunk enters the list (one 7-instruction chain for the whole warp).
            // If for SOME lane a second chunk also beats its fourth-best key, both go in and the third smallest is
            // looked at; only when that one qualifies too -- always while a list is still filling, rarely afterwards --
            // do all eight keys run through the chain.
            if (__any_sync(0xffffffffu, gkf < key[3])) {
                key_insert(key, gkf);
                const float m2 = fminf(fminf(fminf(fmaxf(l01, l23), h01), h23), fminf(fminf(hi0, hi1), fminf(hi2, hi3)));
                const float m2f = __uint_as_float((__float_as_uint(m2) & keep) | idb);
                if (__any_sync(0xffffffffu, m2f < key[3])) {
                    key_insert(key, m2f);
                    float m3 = __uint_as_float(KEY_BIG);
#pragma unroll
                    for (int c = 0; c < 8; c++) m3 = fminf(m3, ck[c] > m2 ? ck[c] : __uint_as_float(KEY_BIG));
                    const float m3f = __uint_as_float((__float_as_uint(m3) & keep) | idb);
                    if (__any_sync(0xffffffffu, m3f < key[3])) {
#pragma unroll
                        for (int c = 0; c < 8; c++) {
                            const float x = __uint_as_float((__float_as_uint(ck[c]) & keep) | idb);
                            key_insert(key, ck[c] > m2 ? x : __uint_as_float(KEY_BIG));
                        }
                    }
                }
            }
#endif
            const bool block_end = (tb_next == 0 || n + 1 == n_tiles);   // last tile of this query block in the span
            if (++tw == KEY_WINDOW || block_end) {
                key_flush(key, abs_id, chunk_base);
                tw = 0;
                chunk_base = tb_next * (BN / CHUNK) + csub * 8;
            }
            if (block_end) {   // publish
                const int slot = (unit - part.cta_of((long long)qb * ntb)) * SUBSLOTS + csub;
                const size_t o = ((size_t)(qb * qmul + rank) * BM + row) * slots + slot;
                float pv[NCAND];
#pragma unroll
                for (int i = 0; i < NCAND; i++)
                    pv[i] = abs_id[i] < 0 ? __int_as_float(0x7f800000) : __uint_as_float(__float_as_uint(key[i]) & ~KEY_ID_MASK);
                cand_v[o] = make_float4(pv[0], pv[1], pv[2], pv[3]);
                cand_id[o] = make_int4(abs_id[0], abs_id[1], abs_id[2], abs_id[3]);
                key_reset(key, abs_id);
                qb++;
            }
            tb = tb_next;
        }
    }

    if (warp == EPI_WARP0 && lane == 0) TC_TRACE(6);   // epilogue of the first epilogue warp done
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) { TC_TRACE(7); }
#ifdef SBA_TC_TRACE
    if (threadIdx.x == 0) g_tc_trace[(blockIdx.x % 148) * 16 + 8] = (unsigned long long)n_tiles;
#endif
    if (PAIR) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair's MMAs and signals may still touch it
    if (warp == 1) {
        tcgen05_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---- 3. exact re-rank ---------------------------------------------------------------------------------------
// The bound test and its margin.  For a train row j outside the evaluated chunks the tensor pass guarantees
//   approx_j >= B        (B = fourth best chunk key of the row, id bits cleared),
// and the row is final iff no such j can beat the exact second best:  d1^2 < B + |a|^2 - delta,  where delta must
// cover  |approx_j - (dsq_j - |a|^2)|  with dsq_j the fp32 value cv::BFMatcher itself would compute.  With A = |a|,
// Bn = |b_j|, u = 2^-24, and "scale" = |a|^2 + max_j |b_j|^2 >= A^2 + Bn^2 >= 2 A Bn:
//   (1) bf16 split: x = hi + lo + r with |r_k| <= 2^-18 |x_k|; the dropped products lo.lo, r.b, a.r are each
//       <= 2^-18 |a_k b_k|, so |a.b - (hi.hi + hi.lo + lo.hi)| <= 3 * 2^-18 A Bn = 1.14e-5 A Bn;
//   (2) the 192 bf16 products are exact in fp32; their accumulation in the tensor core is at worst a sequential
//       truncating fp32 sum: <= 192 * 2^-23 A Bn = 2.3e-5 A Bn (observed: an order of magnitude less);
//   (3) v = fma(-2, S, |b|^2) rounds once (6e-8 scale); |a|^2 and |b|^2 are 64-term fp32 sums (<= 2e-6 of themselves);
//   (4) OpenCV's dsq: every (a_k - b_k) and its square round once, 16 accumulators of 4 terms, 4 + 2 combining adds:
//       <= 10 u dsq <= 6e-7 (A + Bn)^2 <= 1.2e-6 scale.
//   approx = |b|^2 - 2 S, so (1) and (2) enter twice: 2 (1.14e-5 + 2.3e-5) A Bn <= 3.44e-5 scale; with (3) and (4):
//   delta <= 3.8e-5 scale  ->  DELTA_COEF = 4e-5.  The largest error actually seen is reported by sba_match_last_stats
//   (4-5e-6 on random-sign unit rows, 1.3e-5 on all-positive rows; tests/test_gpu_matcher.py holds both).
//   (5) candidate keys: clearing 9 mantissa bits of a chunk minimum and of B moves each by < 2^-14 of its magnitude,
//       and keys that differ only in their id bits may be ordered either way: 3 * 2^-14 |B| = KEY_TRUNC_REL |B| on top.
// 16 threads per query row (8 rows per 128-thread CTA).
//   a) the row's candidate lists (one per span slot) are merged by rank counting: thread e holds up to
//      RR_PER_THREAD entries, every entry's rank = number of entries that sort before it (value, then
//      position), obtained with width-16 shuffles -- no divergent branches.  Ranks 0..3 are the merged
//      best chunks (8 train rows each).
//   b) phase 0: the 16 train rows of the two best chunks are evaluated EXACTLY, one per thread: the 256-byte rows are
//      fetched cooperatively, parked in shared memory and combined with the query row in OpenCV's order
//      (l2sqr_opencv); top-2 by (distance, index) over the 16 threads.  Every other train row has an approximate
//      value >= the third best chunk key, so the row is final if its exact second best clears that bound by the margin.
//   c) phase 1, only for rows that are not final yet: the same for chunks 2 and 3, bound = the fourth best key.
//      Rows that still fail go to the exact fallback.
template <int D> struct RR {                       // 64-d: 8 query rows per 128-thread CTA; 128-d: 4 per 64-thread CTA (shared memory)
    static constexpr int THREADS = D == 64 ? 128 : 64;
    static constexpr int ROWS = THREADS / 16;
};
constexpr int RR_PER_THREAD = 4;   // up to 16*4 = 64 candidate entries (16 slots) merged by shuffles; more -> serial path
static_assert(CHUNK == 8 && NCAND == 4, "the re-rank evaluates two chunks of 8 rows per phase with 16 threads");

template <int D>
__global__ void __launch_bounds__(RR<D>::THREADS, D == 64 ? 6 : 4)   // 37 KB of shared memory per CTA: six fit an SM
tc_rerank_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, const float* __restrict__ na,
                 const float* __restrict__ nb_max, Partition part, const float4* __restrict__ cand_v, const int4* __restrict__ cand_id,
                 int slots, Top2* __restrict__ top, int* __restrict__ fb_list, int* __restrict__ fb_count, float* __restrict__ dbg_max_err,
                 float delta_coef, const float* __restrict__ na_max, const int2* __restrict__ qb_span)
{
    constexpr int RR_ROWS = RR<D>::ROWS, V4 = D / 64;   // V4: float4 pieces of a row per thread of the 16-thread group
    __shared__ __align__(16) float qs[RR_ROWS][D];
    __shared__ __align__(16) float ts[RR_ROWS][16][D + 4];   // the candidate rows of every query of the CTA (one phase)
    __shared__ float win_v[RR_ROWS][NCAND];
    __shared__ int win_id[RR_ROWS][NCAND];
    const int tid = threadIdx.x, grp = tid >> 4, e = tid & 15;
    const int row = blockIdx.x * RR_ROWS + grp;
    const bool active = row < nq;
    const int r = active ? row : nq - 1;
    const float INF = __int_as_float(0x7f800000);
    pdl_trigger();
    pdl_wait();
#ifdef SBA_TC_TRACE
    if (threadIdx.x == 0) atomicMin(&g_tc_edge[1], gtime_early());
#endif

    // a) merge candidate lists.  Everything this row needs from global memory before the exact phase is requested up front: the
    //    span of CTAs that hold lists of its query block (a per-geometry table built on the host; inverting the partition here --
    //    two searches over 64-bit divisions per thread -- was a third of this kernel's instructions), the list entries this thread
    //    owns, and its piece of the query row.
    const int qb = r / BM;
    const int2 sp = __ldg(qb_span + qb);
    const int c_first = sp.x, c_last = sp.y;
    const int n_ent = (c_last - c_first + 1) * SUBSLOTS * NCAND;   // entries are (slot, k) pairs, contiguous in memory
    const float* ev = reinterpret_cast<const float*>(cand_v + (size_t)r * slots);
    const int* ei = reinterpret_cast<const int*>(cand_id + (size_t)r * slots);
    float v[RR_PER_THREAD];
    int id[RR_PER_THREAD];
#pragma unroll
    for (int k = 0; k < RR_PER_THREAD; k++) {
        const int pos = e + 16 * k;
        const bool in = pos < n_ent && n_ent <= 16 * RR_PER_THREAD;
        v[k] = in ? __ldg(ev + pos) : INF;
        id[k] = in ? __ldg(ei + pos) : -1;
    }
    // query rows of this CTA -> shared (coalesced 128-bit loads: 8 rows x 16 float4); a row is written and read by its own
    // 16-thread group only, so a warp-level barrier orders it
#pragma unroll
    for (int vv = 0; vv < V4; vv++) reinterpret_cast<float4*>(&qs[grp][0])[e + 16 * vv] = __ldg(reinterpret_cast<const float4*>(q + (size_t)r * D) + e + 16 * vv);
    if (e < NCAND) { win_v[grp][e] = INF; win_id[grp][e] = -1; }
    __syncwarp();

    if (n_ent <= 16 * RR_PER_THREAD) {
        // Four selection rounds: every thread offers the best of the (<= 4) entries it holds, a width-16 xor butterfly finds the
        // smallest (value, then position -- the same order the lists' slots have in memory), its owner retires it.
        // ~150 instructions per thread; counting ranks over all pairs of entries cost ~800.
#pragma unroll
        for (int round = 0; round < NCAND; round++) {
            float bv = v[0];
            int bk = 0;
#pragma unroll
            for (int k = 1; k < RR_PER_THREAD; k++)
                if (v[k] < bv) { bv = v[k]; bk = k; }            // strict: the lower position wins ties inside a thread
            int bpos = e + 16 * bk;
            float wv = bv;
            int wpos = bpos;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
                const int opos = __shfl_xor_sync(0xffffffffu, wpos, o);
                if (ov < wv || (ov == wv && opos < wpos)) { wv = ov; wpos = opos; }
            }
            // every thread of the group now knows the winner; its owner publishes the id and retires the entry
            if (wpos == bpos && wv < INF) {
                int wid = id[0];
#pragma unroll
                for (int k = 1; k < RR_PER_THREAD; k++) wid = (bk == k) ? id[k] : wid;
                if (wid >= 0) { win_v[grp][round] = wv; win_id[grp][round] = wid; }
#pragma unroll
                for (int k = 0; k < RR_PER_THREAD; k++) v[k] = (bk == k) ? INF : v[k];
            }
        }
    } else if (e == 0) {
        Cand c;
        cand_reset(c);
        for (int pos = 0; pos < n_ent; pos++) cand_insert(c, ev[pos], ei[pos]);
#pragma unroll
        for (int k = 0; k < NCAND; k++) { win_v[grp][k] = c.v[k]; win_id[grp][k] = c.id[k]; }
    }
    __syncwarp();

    const float na_r = na[r];
    const float scale = na_r + *nb_max;
    // fp16 filter only (na_max != NULL): plain fp16 rows are as good as the analysis assumes while the data sits inside fp16's
    // range -- elements below 65504, and norms not so small that whole sets turn subnormal.  Outside that window no row is
    // trusted: everything goes through the exact fallback (slow, never wrong).
    const bool range_ok = na_max == nullptr || (*na_max < 1e9f && *nb_max < 1e9f && fmaxf(*na_max, *nb_max) > 1e-7f);
    Top2 best = top2_empty();
    bool final_row = false;
#pragma unroll
    for (int phase = 0; phase < 2; phase++) {
        const bool need = !final_row;
        if (phase == 1 && !__any_sync(0xffffffffu, need)) break;   // the common case: every row of the warp was settled by phase 0
        // b) the 16 candidate rows of this phase (chunks 2*phase, 2*phase + 1), fetched COOPERATIVELY -- for each row the 16
        //    threads of the group copy its 256 contiguous bytes as one 16-byte piece each, so a request touches 4 lines per
        //    warp instead of 32 -- straight into shared memory with cp.async (row stride 68 floats: the float4 reads below are
        //    conflict free): all 16 rows are in flight together and no register holds them on the way (staging them through
        //    registers, eight rows at a time, cost two dependent round trips and 32 registers).
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const int cid = win_id[grp][2 * phase + half];
#pragma unroll
            for (int cc = 0; cc < 8; cc++) {
                const int cj = cid * CHUNK + cc;
                const bool ok = need && cid >= 0 && cj < nt;
#pragma unroll
                for (int v = 0; v < V4; v++) {
                    const float* src = ok ? t + (size_t)cj * D + 4 * (e + 16 * v) : t;   // nothing is read when ok is false (zero fill)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(&ts[grp][half * 8 + cc][4 * (e + 16 * v)])), "l"(src),
                                 "r"(ok ? 16 : 0)
                                 : "memory");
                }
            }
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        const int chunk = win_id[grp][2 * phase + (e >> 3)];
        const float chunk_v = win_v[grp][2 * phase + (e >> 3)];
        const int j = chunk * CHUNK + (e & 7);
        const bool valid = need && chunk >= 0 && j < nt;
        Top2 cur = top2_empty();
        float exact_v = INF;
        if (valid) {
            // straight from shared memory as float4 (row stride D + 4 floats: conflict free)
            const float dsq = l2sqr_opencv_v4<D>(reinterpret_cast<const float4*>(&qs[grp][0]), reinterpret_cast<const float4*>(&ts[grp][e][0]));
            cur.d0 = __fsqrt_rn(dsq);
            cur.i0 = j;
            exact_v = dsq - na_r;   // exact value on the scale of the approximate ones
        }
        // diagnostics: |chunk key (approx) - min over its 8 rows (exact)|, relative to the bound scale
        float gmin = exact_v;
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, 1));
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, 2));
        gmin = fminf(gmin, __shfl_xor_sync(0xffffffffu, gmin, 4));
        if (active && need && dbg_max_err && (e & 7) == 0 && chunk >= 0 && gmin < INF && scale > 0.f) {
            // net of the key truncation (< 2^-14 |key|, covered separately by KEY_TRUNC_REL): what DELTA_COEF has to cover
            const float err = fmaxf(0.f, fabsf(gmin - chunk_v) - fabsf(chunk_v) * (1.f / 16384.f)) / scale;
            atomicMax((int*)dbg_max_err, __float_as_int(err));   // non-negative floats order like ints
        }
        // top-2 over the 16 threads of the row, then into the running result
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            Top2 other;
            other.d0 = __shfl_xor_sync(0xffffffffu, cur.d0, o);
            other.d1 = __shfl_xor_sync(0xffffffffu, cur.d1, o);
            other.i0 = __shfl_xor_sync(0xffffffffu, cur.i0, o);
            other.i1 = __shfl_xor_sync(0xffffffffu, cur.i1, o);
            cur = top2_merge(cur, other);
        }
        if (need) {
            best = top2_merge(best, cur);
            // final iff no row outside the evaluated chunks can reach the second best:  d1^2 (rounded up) < B + |a|^2 - delta,
            // B = the best key that was NOT evaluated yet (phase 0: rank 2) or the last evaluated one (phase 1: rank 3)
            const float B = win_v[grp][phase == 0 ? 2 : NCAND - 1];
            const float delta = delta_coef * scale + KEY_TRUNC_REL * fabsf(B);
            const float d1sq_up = best.d1 * best.d1 * (1.f + 5e-7f);
            final_row = range_ok && (!(B < INF) || (best.i1 != KNN_MISSING && d1sq_up < B + na_r - delta));
        }
        __syncwarp();
    }
    if (active && e == 0) {
        if (final_row) top[row] = best;
        else {
            const int k = atomicAdd(fb_count, 1);
            fb_list[k] = row;
            Top2 mark;                       // knn2_finalize_kernel merges the fallback's partial lists for this row
            mark.d0 = mark.d1 = 0.f; mark.i0 = KNN_FALLBACK; mark.i1 = k;
            top[row] = mark;
        }
    }
}

// ---- 4. exact fallback for queued rows --------------------------------------------------------------------------
// A queued row needs an exact scan of the whole train set.  Rows are taken FB_ROWS at a time and the train set is cut
// into S ranges (S chosen on the device from the queue length: fb_splits); a work item = (row group, range).  The CTA
// stages its range in shared memory 64 train rows at a time and every staged chunk is used by all rows of the group,
// so the train set is read from L2 once per row GROUP, not once per row: a few dozen queued rows -- what the fp16
// filter's wider margin produces at 16k x 16k -- cost microseconds, and a few thousand (adversarial norms) still spread
// over the whole chip.  knn2_finalize_kernel merges a row's S partial top-2 lists in range order.
constexpr int FB_THREADS = 256;

template <int D>
__global__ void __launch_bounds__(FB_THREADS)
tc_fallback_kernel(const float* __restrict__ q, const float* __restrict__ t, int nt, const int* __restrict__ fb_list,
                   const int* __restrict__ fb_count, Top2* __restrict__ parts)
{
    constexpr int FB_CHUNK = D == 64 ? 64 : 32;   // train rows staged at a time
    constexpr int FB_PITCH = D + 1;               // padded row pitch: lane l reads row l, conflict free
    constexpr int DIM = D;
    __shared__ float qs[FB_ROWS][DIM];
    __shared__ float ts[FB_CHUNK][FB_PITCH];
    pdl_trigger();
    pdl_wait();
    const int n = *fb_count;
    if (n == 0) return;
    const int S = fb_splits(n, gridDim.x);
    const int span = ((nt + S - 1) / S + FB_CHUNK - 1) / FB_CHUNK * FB_CHUNK;   // whole chunks per range
    const int n_groups = (n + FB_ROWS - 1) / FB_ROWS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ROWS_PER_WARP = FB_ROWS / (FB_THREADS / 32);
    for (int item = blockIdx.x; item < n_groups * S; item += gridDim.x) {
        const int grp = item / S, sp = item - grp * S;
        const int k0 = grp * FB_ROWS, nk = min(FB_ROWS, n - k0);
        const int j0 = sp * span, j1 = min(nt, j0 + span);
        __syncthreads();
        for (int i = threadIdx.x; i < FB_ROWS * DIM; i += FB_THREADS) {
            const int r = i / DIM;
            qs[r][i - r * DIM] = r < nk ? q[(size_t)fb_list[k0 + r] * DIM + (i - r * DIM)] : 0.f;
        }
        Top2 best[ROWS_PER_WARP];
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; r++) best[r] = top2_empty();
        for (int c0 = j0; c0 < j1; c0 += FB_CHUNK) {
            __syncthreads();
            for (int i = threadIdx.x; i < FB_CHUNK * DIM / 4; i += FB_THREADS) {      // coalesced 16-byte loads, scalar stores (padded pitch)
                const int r = i / (DIM / 4), c4 = i - r * (DIM / 4);
                const float4 v = (c0 + r < nt) ? __ldg(reinterpret_cast<const float4*>(t + (size_t)(c0 + r) * DIM) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                ts[r][4 * c4] = v.x; ts[r][4 * c4 + 1] = v.y; ts[r][4 * c4 + 2] = v.z; ts[r][4 * c4 + 3] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < ROWS_PER_WARP; r++) {
                const int qr = warp * ROWS_PER_WARP + r;
                if (qr >= nk) break;                       // warp-uniform
#pragma unroll
                for (int hh = 0; hh < FB_CHUNK / 32; hh++) {
                    const int jr = hh * 32 + lane, j = c0 + jr;
                    if (j < j1) {
                        const float d = __fsqrt_rn(l2sqr_opencv<DIM>(&qs[qr][0], &ts[jr][0]));
                        top2_push_ordered(best[r], d, j);   // a lane sees its train rows in increasing order
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; r++) {
            const int qr = warp * ROWS_PER_WARP + r;
            if (qr >= nk) break;
            Top2 b = best[r];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                Top2 other;
                other.d0 = __shfl_xor_sync(0xffffffffu, b.d0, o);
                other.d1 = __shfl_xor_sync(0xffffffffu, b.d1, o);
                other.i0 = __shfl_xor_sync(0xffffffffu, b.i0, o);
                other.i1 = __shfl_xor_sync(0xffffffffu, b.i1, o);
                b = top2_merge(b, other);
            }
            if (lane == 0) parts[(size_t)(k0 + qr) * S + sp] = b;
        }
    }
}

// ---- host ----------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x row_elems 16-bit elements, row-major; box = 64 columns (128 B) x box_rows, 128B swizzle
static int make_map(CUtensorMap* map, void* base, int rows, int box_rows, int row_elems)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SBA_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_elems * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
        return SBA_ERR_CUDA;
    }
    return SBA_OK;
}

}  // namespace tc

bool knn2_tensor_applicable(int nq, int nt, int dim) { return (dim == 64 || dim == 128) && nq >= 1 && nt >= 1; }

// Heuristic used by SBA_MATCH_AUTO: below ~1M pair distances the exact SIMT kernel's latency wins.
bool knn2_tensor_preferred(int nq, int nt, int dim) { return (dim == 64 || dim == 128) && (long long)nq * nt >= (1ll << 20) && nt >= 1024; }

// Both prepared forms of one descriptor set into caller-owned buffers (sba_descriptors_create): prep [n_pad x 128] bf16 hi|lo,
// prep16 [n_pad x 64] fp16, norm [n_pad] (PAD_NORM on the padding rows), *max_norm = largest real norm.
int knn2_prepare_set(sba_ctx* c, const float* d_raw, int n, int n_pad, __nv_bfloat16* prep, float* norm, float* max_norm, void* prep16)
{
    using namespace tc;
    SBA_CUDA(cudaMemsetAsync(max_norm, 0, sizeof(float), c->stream));
    tc_prep_kernel<64><<<(n_pad * 16 + 255) / 256, 256, 0, c->stream>>>(nullptr, 0, 0, nullptr, nullptr, d_raw, n, n_pad, prep, norm, max_norm, nullptr, nullptr,
                                                                        (__half*)prep16);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

// Per query block the first and last CTA whose span holds tiles of it (the re-rank merges exactly those CTAs' lists).  The table
// depends on the partition only, so it is built on the host once per geometry and cached in the context.
// pair mode: the partition counts CTA pairs and super-blocks of two query blocks; the table still has one entry per query block.
static int span_table(sba_ctx* c, const tc::Partition& part, int n_query_blocks, bool pair, const int2** out)
{
    const auto key = std::make_tuple(pair ? -n_query_blocks : n_query_blocks, part.ntb, part.n_ctas, part.bcost);
    auto it = c->tc_spans.find(key);
    if (it == c->tc_spans.end()) {
        if (c->tc_spans.size() >= 256) {   // a sweep over many sizes: start over rather than grow without bound
            SBA_CUDA(cudaStreamSynchronize(c->stream));
            for (auto& kv : c->tc_spans) cudaFree(kv.second.second);
            c->tc_spans.clear();
        }
        std::vector<int> host((size_t)2 * n_query_blocks);
        for (int qb = 0; qb < n_query_blocks; qb++) {
            const int pb = pair ? qb / 2 : qb;   // the partition's block this query block belongs to
            host[2 * qb] = part.cta_of((long long)pb * part.ntb);
            host[2 * qb + 1] = part.cta_of((long long)(pb + 1) * part.ntb - 1);
        }
        int* dev = nullptr;
        SBA_CUDA(cudaMalloc(&dev, host.size() * sizeof(int)));
        it = c->tc_spans.emplace(key, std::make_pair(std::move(host), dev)).first;
        // the host copy lives in the cache entry, so the asynchronous copy may read it whenever it runs
        SBA_CUDA(cudaMemcpyAsync(dev, it->second.first.data(), it->second.first.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    *out = reinterpret_cast<const int2*>(it->second.second);
    return SBA_OK;
}

// One launch of the distance kernel: single CTAs, or clusters of two (pair mode; `part` then counts pairs).
template <int P>
static int launch_knn(sba_ctx* c, bool pair, const tc::Partition& part, cudaStream_t st, const CUtensorMap& map_a, const CUtensorMap& map_b, const float* d_nb,
                      float4* d_cv, int4* d_ci, int slots)
{
    using namespace tc;
    static bool attr_done[2][64] = {};   // function attributes are per device
    if (!attr_done[pair ? 1 : 0][c->device & 63]) {
        if (pair) SBA_CUDA(cudaFuncSetAttribute(tc_knn_kernel<P, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Scheme<P>::SMEM_BYTES));
        else SBA_CUDA(cudaFuncSetAttribute(tc_knn_kernel<P, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Scheme<P>::SMEM_BYTES));
        attr_done[pair ? 1 : 0][c->device & 63] = true;
    }
    if (!pair) {
        SBA_CUDA(launch_pdl(c->pdl, tc_knn_kernel<P, false>, dim3(part.n_ctas), dim3(THREADS), Scheme<P>::SMEM_BYTES, st, map_a, map_b, d_nb, part, d_cv, d_ci, slots));
        return SBA_OK;
    }
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(2 * part.n_ctas); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = Scheme<P>::SMEM_BYTES; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = c->pdl ? 2 : 1;
    SBA_CUDA(cudaLaunchKernelEx(&cfg, tc_knn_kernel<P, true>, map_a, map_b, d_nb, part, d_cv, d_ci, slots));
    return SBA_OK;
}

// Can this device co-schedule a cluster of two of these CTAs (a whole SM each)?  Asked once per device and scheme; a part (or a
// partition of one) that cannot simply keeps the single-CTA kernel.
static bool pair_launchable(int device, int mode)
{
    using namespace tc;
    static int known[4][64] = {};   // 0 = not asked yet, 1 = yes, -1 = no
    int& k = known[mode & 3][device & 63];
    if (k == 0) {
        cudaLaunchConfig_t cfg = {};
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.gridDim = dim3(2); cfg.blockDim = dim3(THREADS); cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e;
        if (mode == 1) {
            cfg.dynamicSmemBytes = Scheme<1>::SMEM_BYTES;
            e = cudaFuncSetAttribute(tc_knn_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Scheme<1>::SMEM_BYTES);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, tc_knn_kernel<1, true>, &cfg);
        } else if (mode == 2) {
            cfg.dynamicSmemBytes = Scheme<2>::SMEM_BYTES;
            e = cudaFuncSetAttribute(tc_knn_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Scheme<2>::SMEM_BYTES);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, tc_knn_kernel<2, true>, &cfg);
        } else {
            cfg.dynamicSmemBytes = Scheme<3>::SMEM_BYTES;
            e = cudaFuncSetAttribute(tc_knn_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Scheme<3>::SMEM_BYTES);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&n, tc_knn_kernel<3, true>, &cfg);
        }
        if (e != cudaSuccess) cudaGetLastError();   // not an error of the call that asked
        k = (e == cudaSuccess && n >= 1) ? 1 : -1;
    }
    return k > 0;
}

// pq / pt: optional prepared forms of the query / train set (64-d only; then that side is not converted again).
// products: 3 = bf16 split filter, 1 = fp16 filter (see tc::Scheme); 128-d descriptors always take the fp16 filter (mode 2).
int knn2_tensor(sba_ctx* c, const float* d_q, int nq, const float* d_t, int nt, int dim, Top2* d_top, const PreparedSet* pq, const PreparedSet* pt,
                int products)
{
    using namespace tc;
    if (dim != 64 && dim != 128) {
        set_error("tensor-core matcher handles 64- and 128-d descriptors only");
        return SBA_ERR_UNSUPPORTED;
    }
    const int mode = dim == 128 ? 2 : (products == 1 ? 1 : 3);
    if (dim == 128) pq = pt = nullptr;           // prepared sets hold the 64-d forms only
    const bool f16 = mode != 3;
    cudaStream_t st = c->stream;
    const int nqb = (nq + BM - 1) / BM, ntb = (nt + BN - 1) / BN;
    const int nq_pad = pq ? pq->n_pad : nqb * BM, nt_pad = pt ? pt->n_pad : ntb * BN;
    // CTA-pair mode (cta_group::2 MMAs: each B half is read from shared memory once for both SMs' tensor cores -- the kernel is
    // bound by shared-memory bandwidth): the partition then counts pairs and super-blocks of two query blocks
    static const int pair_env = std::getenv("SBA_TC_PAIR") ? atoi(std::getenv("SBA_TC_PAIR")) : 1;   // SBA_TC_PAIR=0: single-CTA kernel everywhere
    const int ctas_avail = c->matcher_ctas > 0 ? std::min(c->matcher_ctas, c->sm_count) : c->sm_count;
    const bool pair = pair_env != 0 && ctas_avail >= 2 && ctas_avail % 2 == 0 && nqb >= 2 &&   // an odd CTA budget is honoured exactly by single CTAs
                      pair_launchable(c->device, mode);
    const int npb = pair ? (nqb + 1) / 2 : nqb;   // blocks the partition counts
    Partition part;
    part.nqb = npb; part.ntb = ntb; part.T = (long long)npb * ntb;
    part.n_ctas = (int)std::min<long long>(pair ? ctas_avail / 2 : ctas_avail, part.T);
    // measured with the trace build at 16k x 16k: spans that cross into a new query block finish ~8 us (5-6 tiles) late;
    // capped so that no span can come out empty
    static const int bcost_max = std::getenv("SBA_TC_BCOST") ? std::max(0, atoi(std::getenv("SBA_TC_BCOST"))) : 5;   // tuning knob (tiles)
    part.bcost = (int)std::min<long long>(bcost_max, part.T / part.n_ctas / 4);
    const int slots = part.max_slots() * SUBSLOTS;
    const int row_bytes = mode == 1 ? 128 : 256;   // 16-bit elements: 64 (fp16), 128 (bf16 hi|lo, or fp16 of a 128-d row)

    // workspace carve-up (one buffer)
    auto align_up = [](size_t v) { return (v + 1023) & ~(size_t)1023; };
    size_t off = 0;
    const size_t o_a = off; off = align_up(off + (size_t)nq_pad * row_bytes);
    const size_t o_b = off; off = align_up(off + (size_t)nt_pad * row_bytes);
    const size_t o_na = off; off = align_up(off + (size_t)nq_pad * 4);
    const size_t o_nb = off; off = align_up(off + (size_t)nt_pad * 4);
    const size_t cand_rows = std::max<size_t>((size_t)nq_pad, (size_t)npb * (pair ? 2 : 1) * BM);   // pair mode: an odd last block still has a (padding) partner
    const size_t o_cv = off; off = align_up(off + cand_rows * slots * 16);
    const size_t o_ci = off; off = align_up(off + cand_rows * slots * 16);
    const size_t o_fl = off; off = align_up(off + (size_t)nq * 4);
    const int fb_grid = 2 * c->sm_count;
    // >= n_rows * fb_splits(n_rows, fb_grid) for every n_rows <= nq: FB_MAX_SPLIT lists per row while the queue is short, then ~fb_grid * FB_ROWS + n_rows
    const size_t fb_parts = (size_t)nq + (size_t)fb_grid * FB_ROWS + (size_t)FB_MAX_SPLIT * std::min<size_t>((size_t)nq, (size_t)fb_grid * FB_ROWS / FB_MAX_SPLIT + 1);
    const size_t o_fp = off; off = align_up(off + fb_parts * sizeof(Top2));
    const size_t o_misc = off; off = align_up(off + 64);
    SBA_TRY(c->scratch[SCR_WORK2].ensure(off, st));
    uint8_t* ws = c->scratch[SCR_WORK2].as<uint8_t>();
    const void* dA = pq ? (f16 ? (const void*)pq->prep16 : (const void*)pq->prep) : (const void*)(ws + o_a);
    const void* dB = pt ? (f16 ? (const void*)pt->prep16 : (const void*)pt->prep) : (const void*)(ws + o_b);
    const float* d_na = pq ? pq->norm : (const float*)(ws + o_na);
    const float* d_nb = pt ? pt->norm : (const float*)(ws + o_nb);
    float4* d_cv = (float4*)(ws + o_cv);
    int4* d_ci = (int4*)(ws + o_ci);
    int* d_fl = (int*)(ws + o_fl);
    Top2* d_fparts = (Top2*)(ws + o_fp);
    int* d_fb_count = (int*)(ws + o_misc);
    const int2* d_span = nullptr;
    SBA_TRY(span_table(c, part, nqb, pair, &d_span));
    const float* d_nbmax = pt ? pt->max_norm : (const float*)(ws + o_misc + 4);
    float* d_dbg = (float*)(ws + o_misc + 8);
    const float* d_namax = pq ? pq->max_norm : (const float*)(ws + o_misc + 12);

    SBA_CUDA(cudaMemsetAsync(ws + o_misc, 0, 64, st));
    if (!pq || !pt) {   // convert whichever side arrives as plain fp32 rows (a side that is prepared counts zero rows here)
        const int rows_a = pq ? 0 : nq_pad, rows_b = pt ? 0 : nt_pad;
        __nv_bfloat16* bA = f16 ? nullptr : (__nv_bfloat16*)(ws + o_a);
        __nv_bfloat16* bB = f16 ? nullptr : (__nv_bfloat16*)(ws + o_b);
        __half* hA = f16 ? (__half*)(ws + o_a) : nullptr;
        __half* hB = f16 ? (__half*)(ws + o_b) : nullptr;
        float *na_w = (float*)(ws + o_na), *nb_w = (float*)(ws + o_nb), *nbm = (float*)(ws + o_misc + 4), *nam = (float*)(ws + o_misc + 12);   // maxima zeroed above
        if (dim == 64)
            tc_prep_kernel<64><<<((rows_a + rows_b) * 16 + 255) / 256, 256, 0, st>>>(d_q, nq, rows_a, bA, na_w, d_t, nt, rows_b, bB, nb_w, nbm, nam, hA, hB);
        else
            tc_prep_kernel<128><<<((rows_a + rows_b) * 32 + 255) / 256, 256, 0, st>>>(d_q, nq, rows_a, bA, na_w, d_t, nt, rows_b, bB, nb_w, nbm, nam, hA, hB);
        SBA_LAUNCHED(c);
    }

    CUtensorMap map_a, map_b;
    SBA_TRY(make_map(&map_a, (void*)dA, nq_pad, BM, row_bytes / 2));
    SBA_TRY(make_map(&map_b, (void*)dB, nt_pad, pair ? BN / 2 : BN, row_bytes / 2));   // pair mode: each CTA loads half a tile
    prof_begin(c, SBA_KERNEL_MATCH);
    if (mode == 1) SBA_TRY(launch_knn<1>(c, pair, part, st, map_a, map_b, d_nb, d_cv, d_ci, slots));
    else if (mode == 2) SBA_TRY(launch_knn<2>(c, pair, part, st, map_a, map_b, d_nb, d_cv, d_ci, slots));
    else SBA_TRY(launch_knn<3>(c, pair, part, st, map_a, map_b, d_nb, d_cv, d_ci, slots));
    prof_end(c, SBA_KERNEL_MATCH);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());

    const float dcoef = f16 ? DELTA_COEF_FP16 : DELTA_COEF;
    const float* namax = f16 ? d_namax : nullptr;
    if (dim == 64) {
        SBA_CUDA(launch_pdl(c->pdl, tc_rerank_kernel<64>, dim3((nq + RR<64>::ROWS - 1) / RR<64>::ROWS), dim3(RR<64>::THREADS), 0, st, d_q, nq, d_t, nt, d_na, d_nbmax, part,
                            d_cv, d_ci, slots, d_top, d_fl, d_fb_count, d_dbg, dcoef, namax, d_span));
        SBA_LAUNCHED(c);
        SBA_CUDA(launch_pdl(c->pdl || c->pdl_small, tc_fallback_kernel<64>, dim3(fb_grid), dim3(FB_THREADS), 0, st, d_q, d_t, nt, d_fl, d_fb_count, d_fparts));
    } else {
        SBA_CUDA(launch_pdl(c->pdl, tc_rerank_kernel<128>, dim3((nq + RR<128>::ROWS - 1) / RR<128>::ROWS), dim3(RR<128>::THREADS), 0, st, d_q, nq, d_t, nt, d_na, d_nbmax,
                            part, d_cv, d_ci, slots, d_top, d_fl, d_fb_count, d_dbg, dcoef, namax, d_span));
        SBA_LAUNCHED(c);
        SBA_CUDA(launch_pdl(c->pdl || c->pdl_small, tc_fallback_kernel<128>, dim3(fb_grid), dim3(FB_THREADS), 0, st, d_q, d_t, nt, d_fl, d_fb_count, d_fparts));
    }
    SBA_LAUNCHED(c);
    c->fb_parts = d_fparts; c->fb_count = d_fb_count; c->fb_grid = fb_grid;   // merged per row by knn2_finalize_kernel
    SBA_CUDA(cudaGetLastError());
    // diagnostics (the three words at d_fb_count) reach the pinned mailbox through knn2_finalize_kernel; sba_match_last_stats reads
    // them after a synchronise
    c->match_stats.n_tiles = (int)((long long)nqb * ntb);                 // 256 x 128 tiles, whatever the partition counts
    c->match_stats.n_ctas = pair ? 2 * part.n_ctas : part.n_ctas;
    c->match_stats.n_fallback_rows = -1;  // resolved lazily from the mailbox
    return SBA_OK;
}

}  // namespace sba

#ifdef SBA_TC_TRACE
extern "C" int sba_tc_trace_read(unsigned long long* out /* [148 x 16] */)
{
    return cudaMemcpyFromSymbol(out, sba::tc::g_tc_trace, sizeof(unsigned long long) * 148 * 16) == cudaSuccess ? 0 : -1;
}
// edges[2]: end of the prep kernel, start of the re-rank kernel (ns); reset = 1 re-arms the min/max cells before a call
extern "C" int sba_tc_trace_edges(unsigned long long* edges, int reset)
{
    if (reset) {
        const unsigned long long init[2] = {0ull, ~0ull};
        return cudaMemcpyToSymbol(sba::tc::g_tc_edge, init, sizeof(init)) == cudaSuccess ? 0 : -1;
    }
    return cudaMemcpyFromSymbol(edges, sba::tc::g_tc_edge, sizeof(unsigned long long) * 2) == cudaSuccess ? 0 : -1;
}
#endif
