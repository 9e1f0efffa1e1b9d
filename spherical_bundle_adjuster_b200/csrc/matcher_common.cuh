// matcher_common.cuh -- pieces shared by the exact SIMT matcher and the tensor-core matcher.
#pragma once
#include <cuda_bf16.h>
#include <climits>

#include "common.cuh"

namespace sba {

constexpr int KNN_MISSING = INT_MAX;  // internal "no neighbour" index (sorts after every real index)
constexpr int KNN_FALLBACK = INT_MAX - 1;   // marker in a Top2's i0: the row was queued for the exact fallback, i1 = its queue position
constexpr int FB_MAX_SPLIT = 64;
constexpr int FB_ROWS = 32;     // queued rows evaluated together against one staged range of the train set

// The exact scan of the queued rows is cut into this many train ranges (chosen on the device from the queue length):
// work items = ranges x groups of FB_ROWS rows, enough of them to fill `grid` CTAs whatever the queue length.
__host__ __device__ inline int fb_splits(int n_rows, int grid)
{
    return n_rows <= 0 ? 1 : max(1, min(FB_MAX_SPLIT, (grid * FB_ROWS + n_rows - 1) / n_rows));
}

struct Top2 {
    float d0, d1;
    int i0, i1;
};

__device__ inline Top2 top2_empty()
{
    Top2 t;
    t.d0 = t.d1 = __int_as_float(0x7f800000);
    t.i0 = t.i1 = KNN_MISSING;
    return t;
}

// (distance, index) lexicographic order: OpenCV keeps the lower train index among equal distances.
__device__ inline bool knn_less(float d, int i, float e, int j) { return d < e || (d == e && i < j); }

// Insert a candidate that is visited in increasing index order within one thread.
__device__ inline void top2_push_ordered(Top2& t, float d, int j)
{
    if (d < t.d0) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = j; }
    else if (d < t.d1) { t.d1 = d; t.i1 = j; }
}

// Insert a candidate in arbitrary index order.
__device__ inline void top2_push(Top2& t, float d, int j)
{
    if (knn_less(d, j, t.d0, t.i0)) { t.d1 = t.d0; t.i1 = t.i0; t.d0 = d; t.i0 = j; }
    else if (knn_less(d, j, t.d1, t.i1)) { t.d1 = d; t.i1 = j; }
}

__device__ inline Top2 top2_merge(const Top2& a, const Top2& b)
{
    Top2 r;
    if (knn_less(b.d0, b.i0, a.d0, a.i0)) {
        r.d0 = b.d0; r.i0 = b.i0;
        if (knn_less(a.d0, a.i0, b.d1, b.i1)) { r.d1 = a.d0; r.i1 = a.i0; } else { r.d1 = b.d1; r.i1 = b.i1; }
    } else {
        r.d0 = a.d0; r.i0 = a.i0;
        if (knn_less(b.d0, b.i0, a.d1, a.i1)) { r.d1 = b.d0; r.i1 = b.i0; } else { r.d1 = a.d1; r.i1 = a.i1; }
    }
    return r;
}

// Squared L2 distance of two fp32 rows in OpenCV's normL2Sqr_ order (SSE baseline): element k of
// every 16-chunk goes to accumulator k%16 with a separate multiply and add; lane-wise
// ((d0+d1)+d2)+d3 over the four accumulator vectors; horizontal (s0+s2)+(s1+s3).
// DIM % 16 == 0.  `a`, `b` may point to global or shared memory.
template <int DIM>
__device__ inline float l2sqr_opencv(const float* __restrict__ a, const float* __restrict__ b)
{
    float s[4];
#pragma unroll
    for (int lane = 0; lane < 4; lane++) {
        float S = 0.f;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < DIM / 16; c++) {
                int k = 16 * c + 4 * u + lane;
                float t = __fsub_rn(a[k], b[k]);
                acc = __fadd_rn(acc, __fmul_rn(t, t));
            }
            S = (u == 0) ? acc : __fadd_rn(S, acc);
        }
        s[lane] = S;
    }
    return __fadd_rn(__fadd_rn(s[0], s[2]), __fadd_rn(s[1], s[3]));
}

// The same value from 16-byte aligned rows read as float4 (shared memory: LDS.128), streaming: element 16c + 4u + l is
// component l of piece 4c + u and goes to accumulator (u, l); each accumulator sees its terms in increasing c, as above.
template <int DIM>
__device__ inline float l2sqr_opencv_v4(const float4* __restrict__ a, const float4* __restrict__ b)
{
    float acc[4][4];
#pragma unroll
    for (int c = 0; c < DIM / 16; c++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const float4 x = a[4 * c + u], y = b[4 * c + u];
            const float t0 = __fsub_rn(x.x, y.x), t1 = __fsub_rn(x.y, y.y), t2 = __fsub_rn(x.z, y.z), t3 = __fsub_rn(x.w, y.w);
            if (c == 0) {
                acc[u][0] = __fmul_rn(t0, t0); acc[u][1] = __fmul_rn(t1, t1); acc[u][2] = __fmul_rn(t2, t2); acc[u][3] = __fmul_rn(t3, t3);
            } else {
                acc[u][0] = __fadd_rn(acc[u][0], __fmul_rn(t0, t0)); acc[u][1] = __fadd_rn(acc[u][1], __fmul_rn(t1, t1));
                acc[u][2] = __fadd_rn(acc[u][2], __fmul_rn(t2, t2)); acc[u][3] = __fadd_rn(acc[u][3], __fmul_rn(t3, t3));
            }
        }
    }
    float s[4];
#pragma unroll
    for (int l = 0; l < 4; l++) s[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][l], acc[1][l]), acc[2][l]), acc[3][l]);
    return __fadd_rn(__fadd_rn(s[0], s[2]), __fadd_rn(s[1], s[3]));
}

// A descriptor set prepared once for the tensor-core matcher (bf16 hi/lo rows, norms, largest norm) -- see
// sba_descriptors_create.  n_pad is a multiple of 256 (serves as query block and as train tiles), pad norms are +inf.
struct PreparedSet {
    const float* raw;               // fp32 rows on the device [n x dim]
    const __nv_bfloat16* prep;      // [n_pad x 128] bf16 hi|lo or NULL when dim != 64
    const void* prep16;             // [n_pad x 64] fp16 rows (or NULL)
    const float* norm;              // [n_pad]
    const float* max_norm;          // device scalar
    int n, n_pad, dim;
};

// Runs after either matcher: per-query top-2 -> optional raw kNN output + ratio-test flags.
int launch_knn_finish(sba_ctx* c, const Top2* d_top2, int nq, float ratio, int32_t* d_query_idx, int32_t* d_train_idx, float* d_dist,
                      int32_t* d_n_matches, int32_t* d_knn_idx, float* d_knn_dist, const Top2* d_fb_parts, const int* d_fb_count, int fb_grid);

}  // namespace sba
