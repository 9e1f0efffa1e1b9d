// matcher_simt.cu -- exact fp32 brute-force kNN(k=2) in OpenCV's arithmetic order, the ratio test and
// the ordered compaction of survivors.  Replaces feature_matcher::match_two_image
// (feature_matcher.cpp:42-59) for SBA_MATCH_SIMT_EXACT, and supplies the finishing stage
// (merge + ratio + compaction) of the tensor-core matcher.
//
// SIMT kernel design: a CTA owns 64 query rows and walks 64-row train tiles.  Both tiles sit in
// shared memory transposed ([k][row]) so that one LDS.128 fetches the k-th component of four
// consecutive rows: 2 LDS.128 feed a 4x4 register tile = 16 pairs x (sub, mul, add).  The inner loop
// visits k accumulator-major (m, 16+m, 32+m, 48+m for m = 0..15) which is exactly the order in which
// OpenCV's sixteen partial sums receive their terms, so only one partial sum per pair is live and the
// fp32 result is bit-identical to cv::BFMatcher's.
#include "matcher_common.cuh"

namespace sba {

constexpr int BQ = 64, BT = 64, SIMT_THREADS = 256;

template <int DIM>
__device__ inline void load_tile_regs(const float* __restrict__ src, int row0, int nrows, float4 (&reg)[DIM / 16], int tid)
{
    // chunk index = tid + 256*e ; row = chunk % 64 (consecutive lanes -> consecutive rows, conflict-free
    // transposed stores), kgroup = chunk / 64
#pragma unroll
    for (int e = 0; e < DIM / 16; e++) {
        int chunk = tid + SIMT_THREADS * e;
        int row = chunk & 63, kg = chunk >> 6;
        int gr = row0 + row;
        if (gr < nrows) reg[e] = __ldg(reinterpret_cast<const float4*>(src + (size_t)gr * DIM) + kg);
        else reg[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

template <int DIM>
__device__ inline void store_tile_regs(float* __restrict__ sm, const float4 (&reg)[DIM / 16], int tid)
{
#pragma unroll
    for (int e = 0; e < DIM / 16; e++) {
        int chunk = tid + SIMT_THREADS * e;
        int row = chunk & 63, kg = chunk >> 6;
        sm[(4 * kg + 0) * 64 + row] = reg[e].x;
        sm[(4 * kg + 1) * 64 + row] = reg[e].y;
        sm[(4 * kg + 2) * 64 + row] = reg[e].z;
        sm[(4 * kg + 3) * 64 + row] = reg[e].w;
    }
}

// grid (ceil(nq/64), nsplit).  Split s scans train tiles [tile_begin, tile_end) of its share.
template <int DIM>
__global__ void __launch_bounds__(SIMT_THREADS) knn2_simt_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt,
                                                                 int tiles_per_split, Top2* __restrict__ out /* [nsplit][nq] */)
{
    extern __shared__ __align__(16) float smem[];
    float* qs = smem;             // [DIM][64]
    float* ts = smem + DIM * 64;  // [DIM][64]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * BQ;
    const int n_tiles = (nt + BT - 1) / BT;
    const int tile_begin = blockIdx.y * tiles_per_split;
    const int tile_end = min(n_tiles, tile_begin + tiles_per_split);

    float4 reg[DIM / 16];
    load_tile_regs<DIM>(q, q0, nq, reg, tid);
    store_tile_regs<DIM>(qs, reg, tid);

    Top2 best[4];
#pragma unroll
    for (int a = 0; a < 4; a++) best[a] = top2_empty();

    if (tile_begin < tile_end) load_tile_regs<DIM>(t, tile_begin * BT, nt, reg, tid);
    for (int tile = tile_begin; tile < tile_end; tile++) {
        __syncthreads();  // previous tile fully consumed (and qs visible on the first pass)
        store_tile_regs<DIM>(ts, reg, tid);
        __syncthreads();
        if (tile + 1 < tile_end) load_tile_regs<DIM>(t, (tile + 1) * BT, nt, reg, tid);

        float p02[16], p13[16];
#pragma unroll
        for (int lane = 0; lane < 4; lane++) {
            float S[16];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                float acc[16];
#pragma unroll
                for (int c = 0; c < DIM / 16; c++) {
                    const int k = 16 * c + 4 * u + lane;
                    const float4 qa = *reinterpret_cast<const float4*>(qs + k * 64 + ty * 4);
                    const float4 tb = *reinterpret_cast<const float4*>(ts + k * 64 + tx * 4);
                    const float qv[4] = {qa.x, qa.y, qa.z, qa.w};
                    const float tv[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                    for (int a = 0; a < 4; a++)
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            float d = __fsub_rn(qv[a], tv[b]);
                            float pr = __fmul_rn(d, d);
                            acc[a * 4 + b] = (c == 0) ? pr : __fadd_rn(acc[a * 4 + b], pr);  // 0 + x == x exactly
                        }
                }
#pragma unroll
                for (int e = 0; e < 16; e++) S[e] = (u == 0) ? acc[e] : __fadd_rn(S[e], acc[e]);
            }
#pragma unroll
            for (int e = 0; e < 16; e++) {
                if (lane == 0) p02[e] = S[e];
                else if (lane == 1) p13[e] = S[e];
                else if (lane == 2) p02[e] = __fadd_rn(p02[e], S[e]);
                else p13[e] = __fadd_rn(p13[e], S[e]);
            }
        }
        const int tbase = tile * BT + tx * 4;
#pragma unroll
        for (int b = 0; b < 4; b++) {
            const int j = tbase + b;
            if (j < nt) {
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    float d = __fsqrt_rn(__fadd_rn(p02[a * 4 + b], p13[a * 4 + b]));
                    top2_push_ordered(best[a], d, j);
                }
            }
        }
    }

    // merge the 16 tx lanes that share a query (lanes 0-15 / 16-31 of each warp)
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            Top2 other;
            other.d0 = __shfl_xor_sync(0xffffffffu, best[a].d0, o);
            other.d1 = __shfl_xor_sync(0xffffffffu, best[a].d1, o);
            other.i0 = __shfl_xor_sync(0xffffffffu, best[a].i0, o);
            other.i1 = __shfl_xor_sync(0xffffffffu, best[a].i1, o);
            best[a] = top2_merge(best[a], other);
        }
        const int qi = q0 + ty * 4 + a;
        if (tx == 0 && qi < nq) out[(size_t)blockIdx.y * nq + qi] = best[a];
    }
}

// Merge the per-split partial results of one query.
__global__ void knn2_merge_splits_kernel(const Top2* __restrict__ parts, int nq, int nsplit, Top2* __restrict__ out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    Top2 r = parts[i];
    for (int s = 1; s < nsplit; s++) r = top2_merge(r, parts[(size_t)s * nq + i]);
    out[i] = r;
}

// Ratio test (feature_matcher.cpp:47-56) + ordered compaction of the survivors in ONE launch (knn2_finalize_kernel):
// FIN_THREADS queries per CTA.  Every CTA publishes its survivor count tagged with the call's epoch, then sums the counts of
// the CTAs before it (spinning only on counts that are not there yet -- all of them are written within a microsecond of
// the launch, so this is one hop, not a chain), scans its own flags and scatters: survivors stay in ascending query
// order.  Rows the tensor path queued for the exact fallback arrive as a marker in `top` (i0 == KNN_FALLBACK, i1 =
// position in the queue) and their partial top-2 lists (one per scanned range) are merged here in range order.
#ifndef SBA_FIN_THREADS
#define SBA_FIN_THREADS 1024
#endif
constexpr int FIN_THREADS = SBA_FIN_THREADS;   // queries per CTA (measured: 1024 -> 6.4 us for 16 384 rows under ncu, 256 -> 8.4 us)
constexpr int FIN_EPOCH_SHIFT = 11;   // counts are <= 1024

__device__ inline int keep_flag(const Top2& r, float ratio)
{
    return (r.i0 != KNN_MISSING && r.i1 != KNN_MISSING && r.d0 < __fmul_rn(ratio, r.d1)) ? 1 : 0;
}

__global__ void __launch_bounds__(FIN_THREADS)
knn2_finalize_kernel(const Top2* __restrict__ top, int nq, float ratio, int32_t* __restrict__ knn_idx, float* __restrict__ knn_dist,
                     int32_t* __restrict__ query_idx, int32_t* __restrict__ train_idx, float* __restrict__ dist, int32_t* __restrict__ n_matches,
                     unsigned int* __restrict__ counts /* [-1] = arrival ticket */, unsigned int epoch, const Top2* __restrict__ fb_parts, const int* __restrict__ fb_count,
                     int fb_grid, int32_t* __restrict__ mail /* the context's pinned host mailbox, written directly: [0] = count, [8..10] = tensor-path counters */)
{
    __shared__ int warp_off[32];
    __shared__ int s_base, s_blk;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    pdl_wait();
    // Logical block number = arrival order (a ticket in the slot BEFORE the counts): a block only ever waits for blocks that are
    // already running, whatever order the hardware starts them in.  The last ticket resets the counter for the next call.
    if (tid == 0) {
        const unsigned int t = atomicAdd(counts - 1, 1u);
        if (t == gridDim.x - 1) *(counts - 1) = 0u;
        s_blk = (int)t;
    }
    __syncthreads();
    const int blk = s_blk;
    const int i = blk * FIN_THREADS + tid;
    Top2 r = top2_empty();
    int keep = 0;
    if (i < nq) {
        r = top[i];
        if (fb_parts && r.i0 == KNN_FALLBACK) {
            const int S = fb_splits(*fb_count, fb_grid);
            const Top2* part = fb_parts + (size_t)r.i1 * S;
            r = part[0];
            for (int sp = 1; sp < S; sp++) r = top2_merge(r, part[sp]);
        }
        if (knn_idx) {
            knn_idx[2 * i] = r.i0 != KNN_MISSING ? r.i0 : -1;
            knn_idx[2 * i + 1] = r.i1 != KNN_MISSING ? r.i1 : -1;
        }
        if (knn_dist) {
            knn_dist[2 * i] = r.d0;
            knn_dist[2 * i + 1] = r.d1;
        }
        keep = keep_flag(r, ratio);
    }
    const unsigned ball = __ballot_sync(0xffffffffu, keep);
    const int prefix = __popc(ball & ((1u << lane) - 1));
    if (lane == 0) warp_off[warp] = __popc(ball);
    __syncthreads();
    int total = 0;
    if (warp == 0) {
        const int v = lane < FIN_THREADS / 32 ? warp_off[lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        warp_off[lane] = incl - v;
        total = __shfl_sync(0xffffffffu, incl, 31);
        if (lane == 0) {
            __threadfence();
            ((volatile unsigned int*)counts)[blk] = (epoch << FIN_EPOCH_SHIFT) | (unsigned int)total;
        }
        // counts of the CTAs before this one (this warp strides over them)
        int before = 0;
        for (int b = lane; b < blk; b += 32) {
            unsigned int v2;
            do { v2 = ((volatile unsigned int*)counts)[b]; } while ((v2 >> FIN_EPOCH_SHIFT) != epoch);
            before += (int)(v2 & ((1u << FIN_EPOCH_SHIFT) - 1));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
        if (lane == 0) {
            s_base = before;
            if (blk == (int)gridDim.x - 1) {
                *n_matches = before + total;
                mail[0] = before + total;      // the host reads the mailbox after its synchronise: no copy node on the stream
                if (fb_count) { mail[8] = fb_count[0]; mail[9] = fb_count[1]; mail[10] = fb_count[2]; }
            }
        }
    }
    __syncthreads();
    if (keep) {
        const int pos = s_base + warp_off[warp] + prefix;
        query_idx[pos] = i;
        train_idx[pos] = r.i0;
        dist[pos] = r.d0;
    }
}

int launch_knn_finish(sba_ctx* c, const Top2* d_top2, int nq, float ratio, int32_t* d_query_idx, int32_t* d_train_idx, float* d_dist,
                      int32_t* d_n_matches, int32_t* d_knn_idx, float* d_knn_dist, const Top2* d_fb_parts, const int* d_fb_count, int fb_grid)
{
    const int nblocks = (nq + FIN_THREADS - 1) / FIN_THREADS;
    // the epoch-tagged count slots live in their own grow-only buffer; a fresh (or regrown) buffer and an epoch wrap start from zeros
    const size_t cap_before = c->scratch[SCR_FIN_COUNTS].cap;
    SBA_TRY(c->scratch[SCR_FIN_COUNTS].ensure((size_t)(nblocks + 1) * sizeof(unsigned int), c->stream));   // counts + the ticket
    c->fin_epoch = (c->fin_epoch + 1) & ((1u << (32 - FIN_EPOCH_SHIFT)) - 1);
    if (c->scratch[SCR_FIN_COUNTS].cap != cap_before || c->fin_epoch == 0) {
        SBA_CUDA(cudaMemsetAsync(c->scratch[SCR_FIN_COUNTS].p, 0, c->scratch[SCR_FIN_COUNTS].cap, c->stream));
        if (c->fin_epoch == 0) c->fin_epoch = 1;
    }
    SBA_CUDA(launch_pdl(c->pdl || c->pdl_small, knn2_finalize_kernel, dim3(nblocks), dim3(FIN_THREADS), 0, c->stream, d_top2, nq, ratio, d_knn_idx, d_knn_dist, d_query_idx, d_train_idx,
                        d_dist, d_n_matches, c->scratch[SCR_FIN_COUNTS].as<unsigned int>() + 1, c->fin_epoch, d_fb_parts, d_fb_count, fb_grid,
                        (int32_t*)c->pinned_i32));
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

template <int DIM>
static int run_simt(sba_ctx* c, const float* d_q, int nq, const float* d_t, int nt, Top2* d_top)
{
    const int qblocks = (nq + BQ - 1) / BQ;
    const int n_tiles = (nt + BT - 1) / BT;
    // split the train range until the grid has ~3 CTAs per SM (occupancy of this kernel) or each
    // split is down to 4 tiles
    int nsplit = 1;
    const int target = c->sm_count * 3;
    if (qblocks < target) {
        nsplit = std::min((target + qblocks - 1) / qblocks, std::max(1, n_tiles / 4));
        nsplit = std::max(1, std::min(nsplit, 64));
    }
    int tiles_per_split = (n_tiles + nsplit - 1) / nsplit;
    if (tiles_per_split < 1) tiles_per_split = 1;
    nsplit = std::max(1, (n_tiles + tiles_per_split - 1) / tiles_per_split);
    const size_t smem = (size_t)2 * DIM * 64 * sizeof(float);
    SBA_CUDA(cudaFuncSetAttribute(knn2_simt_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Top2* d_parts = d_top;
    if (nsplit > 1) {
        SBA_TRY(c->scratch[SCR_WORK1].ensure((size_t)nsplit * nq * sizeof(Top2), c->stream));
        d_parts = c->scratch[SCR_WORK1].as<Top2>();
    }
    dim3 grid(qblocks, nsplit);
    prof_begin(c, SBA_KERNEL_MATCH);
    knn2_simt_kernel<DIM><<<grid, SIMT_THREADS, smem, c->stream>>>(d_q, nq, d_t, nt, tiles_per_split, d_parts);
    prof_end(c, SBA_KERNEL_MATCH);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    if (nsplit > 1) {
        knn2_merge_splits_kernel<<<(nq + 255) / 256, 256, 0, c->stream>>>(d_parts, nq, nsplit, d_top);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
    }
    return SBA_OK;
}

// defined in matcher_tc.cu
int knn2_tensor(sba_ctx* c, const float* d_q, int nq, const float* d_t, int nt, int dim, Top2* d_top, const PreparedSet* pq, const PreparedSet* pt, int products);
int knn2_prepare_set(sba_ctx* c, const float* d_raw, int n, int n_pad, __nv_bfloat16* prep, float* norm, float* max_norm, void* prep16);
bool knn2_tensor_applicable(int nq, int nt, int dim);
bool knn2_tensor_preferred(int nq, int nt, int dim);

__global__ void fill_empty_top2_kernel(Top2* top, int nq)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq) top[i] = top2_empty();
}

}  // namespace sba

using namespace sba;

// A descriptor set kept on the device in both forms (sba_descriptors_create).
struct sba_descriptors {
    int device = 0;
    float* raw = nullptr;
    __nv_bfloat16* prep = nullptr;
    void* prep16 = nullptr;
    float* norm = nullptr;          // [n_pad] + the maximum in the last slot
    sba::PreparedSet set{};
};

extern "C" {

// Shared by the two entry points: q/t are fp32 rows where `mem` says, unless the side comes prepared (then its rows
// already live on the device).
static int knn2_ratio_impl(sba_ctx* c, const float* q, int nq, const float* t, int nt, int dim, float ratio, int32_t* query_idx, int32_t* train_idx,
                           float* dist, int32_t* n_matches, int32_t* knn_idx, float* knn_dist, int mem, int algo, const PreparedSet* pq,
                           const PreparedSet* pt)
{
    SBA_CHECK_ARG(c && nq >= 0 && nt >= 0 && n_matches);
    SBA_CHECK_ARG(nq == 0 || (q && query_idx && train_idx && dist));
    SBA_CHECK_ARG(nt == 0 || t);
    if (dim != 64 && dim != 128) {
        set_error("descriptor dimension %d not supported (SURF: 64 or 128)", dim);
        return SBA_ERR_UNSUPPORTED;
    }
    SBA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->match_stats = sba_match_stats{};
    c->fb_parts = nullptr; c->fb_count = nullptr; c->fb_grid = 0;   // set by the tensor path when it queues fallback rows
    if (nq == 0) {
        if (mem == SBA_MEM_HOST) *n_matches = 0;
        else SBA_CUDA(cudaMemsetAsync(n_matches, 0, sizeof(int32_t), st));
        return finish(c, mem);
    }
    const float *d_q = q, *d_t = t;
    if (!pq) SBA_TRY(stage_in(c, q, (size_t)nq * dim, mem, SCR_IN0, &d_q));
    if (!pt) SBA_TRY(stage_in(c, t, (size_t)nt * dim, mem, SCR_IN1, &d_t));
    int32_t *d_qi, *d_ti, *d_n, *d_ki;
    float *d_d, *d_kd;
    SBA_TRY(stage_out(c, query_idx, (size_t)nq, mem, SCR_OUT0, &d_qi));
    SBA_TRY(stage_out(c, train_idx, (size_t)nq, mem, SCR_OUT1, &d_ti));
    SBA_TRY(stage_out(c, dist, (size_t)nq, mem, SCR_OUT2, &d_d));
    SBA_TRY(stage_out(c, knn_idx, (size_t)2 * nq, mem, SCR_OUT3, &d_ki));
    SBA_TRY(stage_out(c, knn_dist, (size_t)2 * nq, mem, SCR_OUT4, &d_kd));
    SBA_TRY(c->scratch[SCR_WORK0].ensure((size_t)nq * sizeof(Top2) + 16, st));
    Top2* d_top = c->scratch[SCR_WORK0].as<Top2>();
    if (mem == SBA_MEM_HOST) {
        SBA_TRY(c->scratch[SCR_WORK5].ensure(sizeof(int32_t), st));
        d_n = c->scratch[SCR_WORK5].as<int32_t>();
    } else {
        d_n = n_matches;
    }

    int use = algo;
    // AUTO takes the bf16-split filter: measured on B200 the single-product fp16 filter saves nothing in the distance kernel (it is
    // bound by its epilogue, not by the tensor pipe) and its wider margin sends ~0.3 % of the rows to the exact fallback
    if (use == SBA_MATCH_AUTO) use = knn2_tensor_preferred(nq, nt, dim) ? SBA_MATCH_TENSOR : SBA_MATCH_SIMT_EXACT;
    if (nt == 0) {
        fill_empty_top2_kernel<<<(nq + 255) / 256, 256, 0, st>>>(d_top, nq);
        SBA_LAUNCHED(c);
    } else if (use == SBA_MATCH_TENSOR || use == SBA_MATCH_TENSOR_FP16) {
        if (!knn2_tensor_applicable(nq, nt, dim)) {
            set_error("tensor-core matcher does not apply to nq=%d nt=%d dim=%d", nq, nt, dim);
            return SBA_ERR_UNSUPPORTED;
        }
        SBA_TRY(knn2_tensor(c, d_q, nq, d_t, nt, dim, d_top, (pq && pq->prep) ? pq : nullptr, (pt && pt->prep) ? pt : nullptr,
                            use == SBA_MATCH_TENSOR ? 3 : 1));
    } else {
        use = SBA_MATCH_SIMT_EXACT;
        if (dim == 64) SBA_TRY(run_simt<64>(c, d_q, nq, d_t, nt, d_top));
        else SBA_TRY(run_simt<128>(c, d_q, nq, d_t, nt, d_top));
    }
    c->match_stats.algo_used = use;
    SBA_TRY(launch_knn_finish(c, d_top, nq, ratio, d_qi, d_ti, d_d, d_n, d_ki, d_kd, (const Top2*)c->fb_parts, c->fb_count, c->fb_grid));
    if (mem == SBA_MEM_HOST) {
        SBA_CUDA(cudaStreamSynchronize(st));
        const int n = c->pinned_i32[0];     // written by knn2_finalize_kernel
        *n_matches = n;
        SBA_TRY(copy_out(c, query_idx, d_qi, (size_t)n, mem));
        SBA_TRY(copy_out(c, train_idx, d_ti, (size_t)n, mem));
        SBA_TRY(copy_out(c, dist, d_d, (size_t)n, mem));
        SBA_TRY(copy_out(c, knn_idx, d_ki, (size_t)2 * nq, mem));
        SBA_TRY(copy_out(c, knn_dist, d_kd, (size_t)2 * nq, mem));
    }
    return finish(c, mem);
}

int sba_knn2_ratio(sba_ctx* c, const float* q, int nq, const float* t, int nt, int dim, float ratio, int32_t* query_idx, int32_t* train_idx,
                   float* dist, int32_t* n_matches, int32_t* knn_idx, float* knn_dist, int mem, int algo)
{
    return knn2_ratio_impl(c, q, nq, t, nt, dim, ratio, query_idx, train_idx, dist, n_matches, knn_idx, knn_dist, mem, algo, nullptr, nullptr);
}

int sba_descriptors_create(sba_ctx* c, const float* desc, int n, int dim, int mem, sba_descriptors** out)
{
    SBA_CHECK_ARG(c && out && n >= 0 && (n == 0 || desc));
    *out = nullptr;
    if (dim != 64 && dim != 128) {
        set_error("descriptor dimension %d not supported (SURF: 64 or 128)", dim);
        return SBA_ERR_UNSUPPORTED;
    }
    SBA_CUDA(cudaSetDevice(c->device));
    sba_descriptors* d = new sba_descriptors();
    d->device = c->device;
    d->set.n = n; d->set.dim = dim; d->set.n_pad = (n + 255) / 256 * 256;
    auto fail = [&](cudaError_t e) { set_error("%s", cudaGetErrorString(e)); sba_descriptors_destroy(d); return SBA_ERR_CUDA; };
    cudaError_t e;
    const size_t raw_bytes = (size_t)std::max(n, 1) * dim * sizeof(float);
    if ((e = cudaMalloc(&d->raw, raw_bytes)) != cudaSuccess) return fail(e);
    if (n > 0 && (e = cudaMemcpyAsync(d->raw, desc, (size_t)n * dim * sizeof(float), mem == SBA_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice,
                                      c->stream)) != cudaSuccess)
        return fail(e);
    d->set.raw = d->raw;
    if (dim == 64 && n > 0) {
        if ((e = cudaMalloc(&d->prep, (size_t)d->set.n_pad * 128 * sizeof(__nv_bfloat16))) != cudaSuccess) return fail(e);
        if ((e = cudaMalloc(&d->norm, (size_t)d->set.n_pad * sizeof(float) + sizeof(float))) != cudaSuccess) return fail(e);
        if ((e = cudaMalloc(&d->prep16, (size_t)d->set.n_pad * 64 * 2)) != cudaSuccess) return fail(e);
        float* max_norm = d->norm + d->set.n_pad;
        int rc = knn2_prepare_set(c, d->raw, n, d->set.n_pad, d->prep, d->norm, max_norm, d->prep16);
        if (rc != SBA_OK) { sba_descriptors_destroy(d); return rc; }
        d->set.prep = d->prep; d->set.norm = d->norm; d->set.max_norm = max_norm; d->set.prep16 = d->prep16;
    }
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return fail(e);   // usable from any context / stream of this device from here on
    *out = d;
    return SBA_OK;
}

int sba_descriptors_destroy(sba_descriptors* d)
{
    if (!d) return SBA_OK;
    cudaSetDevice(d->device);
    cudaDeviceSynchronize();          // a match that reads the set may still be queued on some stream
    if (d->raw) cudaFree(d->raw);
    if (d->prep) cudaFree(d->prep);
    if (d->prep16) cudaFree(d->prep16);
    if (d->norm) cudaFree(d->norm);
    delete d;
    return SBA_OK;
}

int sba_descriptors_count(const sba_descriptors* d) { return d ? d->set.n : 0; }

int sba_knn2_ratio_prepared(sba_ctx* c, const sba_descriptors* query, const sba_descriptors* train, float ratio, int32_t* query_idx,
                            int32_t* train_idx, float* dist, int32_t* n_matches, int32_t* knn_idx, float* knn_dist, int mem, int algo)
{
    SBA_CHECK_ARG(c && query && train);
    if (query->set.dim != train->set.dim || query->device != c->device || train->device != c->device) {
        set_error("prepared descriptor sets must share the dimension and live on the context's device");
        return SBA_ERR_INVALID;
    }
    return knn2_ratio_impl(c, query->set.raw, query->set.n, train->set.raw, train->set.n, query->set.dim, ratio, query_idx, train_idx, dist,
                           n_matches, knn_idx, knn_dist, mem, algo, &query->set, &train->set);
}

int sba_match_last_stats(sba_ctx* c, sba_match_stats* out)
{
    SBA_CHECK_ARG(c && out);
    if (c->match_stats.n_fallback_rows < 0) {
        // tensor path: knn2_finalize_kernel copied the counters into the pinned mailbox
        SBA_CUDA(cudaStreamSynchronize(c->stream));
        c->match_stats.n_fallback_rows = c->pinned_i32[8];
        memcpy(&c->match_stats.max_rel_err, c->pinned_i32 + 10, sizeof(float));
    }
    *out = c->match_stats;
    return SBA_OK;
}

}  // extern "C"
