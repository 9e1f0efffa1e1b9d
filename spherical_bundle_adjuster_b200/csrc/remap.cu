// remap.cu -- equi2cube gather (K1), cube2equi keypoints (K2), pixel->bearing (K4), match gather.
//
// Replaces equi2cube.cpp:12-302, equi2cube_surf.cpp:19-76,:107-113 and
// spherical_bundle_adjuster.cpp:271-298 of the reference.
//
// K1 design (HBM-bound byte gather; no tensor cores, no textures):
//   The source index of an output pixel depends only on the geometry (w, h, cube_size), so it is
//   computed once into an int32 table ("remap plan", cached per geometry in the context) and every
//   frame after that is a pure table-driven gather:  4 B index read + 3 B gathered + 3 B written per
//   output pixel.  A warp owns 128 adjacent strip pixels: lanes walk adjacent pixels when gathering (a load
//   instruction then touches the few source lines one short curve segment crosses) and the warp's 384 output
//   bytes leave as 24 aligned 16-byte stores (details at remap_gather_warp_kernel).
//   The table is built on the device with the reference's fp64 formula.  CUDA's fp64 acos/atan2 are
//   not correctly rounded, so any pixel whose continuous source coordinate lies within 1e-6 of an
//   integer (the only place a last-ulp difference can flip the truncation; ~0.6 % of pixels: axes
//   and diagonals) is re-evaluated on the host with the same glibc libm the reference links and
//   patched in.  The table is therefore bit-identical to the reference's index arithmetic.
#include <climits>
#include <cmath>

#include "common.cuh"
#include "bulk.cuh"
#include "geometry.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace sba {

// ---- plan construction ------------------------------------------------------------------------
__global__ void lut_build_kernel(int cs, int w, int h, int32_t* __restrict__ lut, int32_t* __restrict__ flagged,
                                 int* __restrict__ n_flagged, int flag_cap, int* __restrict__ n_clamped)
{
    int64_t total = (int64_t)cs * 6 * cs;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        int i = (int)(p / (6 * cs));
        int rem = (int)(p - (int64_t)i * 6 * cs);
        int f = rem / cs, j = rem - f * cs;
        double v[3], rf, cf;
        face_cart(f, (double)i, (double)j, (double)cs, v);
        dir_to_erp(v, w, h, &rf, &cf);
        int c;
        lut[p] = clamp_index(rf, cf, w, h, &c);
        if (c) atomicAdd(n_clamped, 1);
        const double tol = 1e-6;
        bool near = fabs(rf - rint(rf)) < tol || fabs(cf - rint(cf)) < tol;
        if (near) {
            int slot = atomicAdd(n_flagged, 1);
            if (slot < flag_cap) flagged[slot] = (int32_t)p;
        }
    }
}

__global__ void lut_patch_kernel(int32_t* __restrict__ lut, const int32_t* __restrict__ where, const int32_t* __restrict__ what, int n)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) lut[where[k]] = what[k];
}

static int32_t host_src_index(int p, int cs, int w, int h, int* clamped)
{
    int i = p / (6 * cs), rem = p - i * 6 * cs, f = rem / cs, j = rem - f * cs;
    double v[3], rf, cf;
    face_cart(f, (double)i, (double)j, (double)cs, v);
    dir_to_erp(v, w, h, &rf, &cf);
    return clamp_index(rf, cf, w, h, clamped);
}

static int get_plan(sba_ctx* c, int w, int h, int cs, RemapPlan** out)
{
    auto key = std::make_tuple(w, h, cs);
    auto it = c->plans.find(key);
    if (it != c->plans.end()) { *out = &it->second; return SBA_OK; }

    int64_t total = (int64_t)cs * 6 * cs;
    SBA_CHECK_ARG(total * 1 < (int64_t)1 << 31 && (int64_t)w * h < (int64_t)1 << 31);
    RemapPlan plan;
    plan.w = w; plan.h = h; plan.cs = cs;
    SBA_CUDA(cudaMalloc(&plan.lut, total * sizeof(int32_t)));
    int flag_cap = (int)total;  // worst case every pixel is flagged
    SBA_TRY(c->scratch[SCR_WORK0].ensure((size_t)flag_cap * sizeof(int32_t), c->stream));
    SBA_TRY(c->scratch[SCR_WORK1].ensure(2 * sizeof(int), c->stream));
    int32_t* d_flag = c->scratch[SCR_WORK0].as<int32_t>();
    int* d_cnt = c->scratch[SCR_WORK1].as<int>();
    SBA_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), c->stream));
    int threads = 256;
    int blocks = (int)std::min<int64_t>(ceil_div64(total, threads), (int64_t)c->sm_count * 8);
    lut_build_kernel<<<blocks, threads, 0, c->stream>>>(cs, w, h, plan.lut, d_flag, d_cnt, flag_cap, d_cnt + 1);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    int cnt[2];
    SBA_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    int nflag = std::min(cnt[0], flag_cap);
    plan.n_patched = nflag;
    if (nflag > 0) {
        std::vector<int32_t> where(nflag), what(nflag);
        SBA_CUDA(cudaMemcpyAsync(where.data(), d_flag, (size_t)nflag * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        SBA_CUDA(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < nflag; k++) what[k] = host_src_index(where[k], cs, w, h, nullptr);
        SBA_TRY(c->scratch[SCR_WORK2].ensure((size_t)nflag * sizeof(int32_t), c->stream));
        int32_t* d_what = c->scratch[SCR_WORK2].as<int32_t>();
        SBA_CUDA(cudaMemcpyAsync(d_what, what.data(), (size_t)nflag * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        lut_patch_kernel<<<(nflag + 255) / 256, 256, 0, c->stream>>>(plan.lut, d_flag, d_what, nflag);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
        SBA_CUDA(cudaStreamSynchronize(c->stream));  // `what` goes out of scope
    }
    plan.n_clamped = cnt[1];
    SBA_TRY(build_tiled_plan(c, plan.lut, cs, 6 * cs, w, h, false, &plan.tiled));
    auto ins = c->plans.emplace(key, plan);
    *out = &ins.first->second;
    return SBA_OK;
}

// ---- the gather -------------------------------------------------------------------------------
// Fast path.  A warp owns 128 adjacent output pixels (384 output bytes), four rounds of 32.
//   gather : in round k lane l fetches pixel 32k+l, so the 32 lanes of one load instruction read ~32 ADJACENT
//            source pixels (100-140 contiguous bytes: one or two L1 wavefronts instead of the four to six a
//            4-pixels-per-lane layout costs).  A pixel's 3 bytes come from the aligned 32-bit word holding its
//            first byte, plus the next word only when it straddles (byte offset 2 or 3), extracted with a
//            funnel shift;
//   store  : pixels go to shared memory one word each (conflict free), each lane reads back ITS four adjacent
//            pixels as one 16-byte load, packs them to 12 bytes, and a second staging row lets lanes 0..23 write
//            the warp's 384 bytes as 24 aligned 16-byte stores.
// Requirements (checked by the launcher): image bases 4-byte aligned, table and output bases 16-byte aligned,
// P*3 % 16 == 0; pixels past the last full 128-pixel group of an image go to remap_gather1_kernel.
// MASKED: table entries < 0 mean "no source pixel" and produce 0 (spherical_surf's bounds check).
template <bool MASKED>
__global__ void __launch_bounds__(256)
remap_gather_warp_kernel(const uint8_t* __restrict__ erp, const int32_t* __restrict__ lut, uint8_t* __restrict__ out,
                         int64_t src_bytes_per_image, int64_t P, int groups128, int n_images,
                         const uint8_t* __restrict__ erp2 = nullptr, uint8_t* __restrict__ out2 = nullptr, int n_first = 0)
{
    // erp2/out2: a second, separately allocated run of images (images >= n_first come from there) -- lets the pair pipeline
    // remap its two ERP images, which live in two buffers, with ONE launch.
    __shared__ __align__(16) uint32_t pix[8][128];
    __shared__ __align__(16) uint32_t stage[8][96];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int total = groups128 * n_images;          // 32-bit group arithmetic (the launcher splits larger batches)
    const int64_t last_word = src_bytes_per_image / 4 - 1;
    for (int g = blockIdx.x * 8 + wib; g < total; g += gridDim.x * 8) {
        int img = g / groups128;
        const int grp = g - img * groups128;
        const bool second = erp2 != nullptr && img >= n_first;
        if (second) img -= n_first;
        const uint32_t* words = reinterpret_cast<const uint32_t*>((second ? erp2 : erp) + (int64_t)img * src_bytes_per_image);
        const int32_t* tab = lut + (int64_t)grp * 128 + lane;
        int idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) idx[k] = __ldg(tab + 32 * k);
        uint32_t w0[4], w1[4], sh[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {          // all loads first (memory-level parallelism)
            const int64_t a = (int64_t)((MASKED && idx[k] < 0) ? 0 : idx[k]) * 3;
            const int64_t wi = a >> 2;
            sh[k] = (uint32_t)(a & 3) * 8;
            w0[k] = __ldg(words + wi);
            w1[k] = (sh[k] > 8) ? __ldg(words + min(wi + 1, last_word)) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t v = __funnelshift_r(w0[k], w1[k], sh[k]) & 0x00FFFFFFu;
            if (MASKED && idx[k] < 0) v = 0u;
            pix[wib][32 * k + lane] = v;
        }
        __syncwarp();
        const uint4 q = *reinterpret_cast<const uint4*>(&pix[wib][4 * lane]);   // pixels 4*lane .. 4*lane+3
        stage[wib][lane * 3 + 0] = q.x | (q.y << 24);
        stage[wib][lane * 3 + 1] = (q.y >> 8) | (q.z << 16);
        stage[wib][lane * 3 + 2] = (q.z >> 16) | (q.w << 8);
        __syncwarp();
        if (lane < 24) {
            const uint4 o = *reinterpret_cast<const uint4*>(&stage[wib][lane * 4]);
            uint8_t* dst = (second ? out2 : out) + ((int64_t)img * P + (int64_t)grp * 128) * 3 + lane * 16;
            *reinterpret_cast<uint4*>(dst) = o;
        }
        __syncwarp();   // both staging rows are rewritten by the next group
    }
}

// ---- tiled variant ------------------------------------------------------------------------------------------
// The direct gather above is bound by the L1 data pipe: a run of 32 output pixels maps onto a short curve that
// crosses ~6 source rows, every row is another 128-byte line, and each line costs a wavefront -- four per load
// instruction, eight load instructions per 128 pixels.  Here a CTA owns a 16 x 128 output tile whose source pixels
// lie in a compact bounding box (median 15 KB at 4K -> cube 960): the box is staged in shared memory by 1-D bulk
// copies, one per source row, issued by warp 0 -- the TMA engine moves the bytes, the LSU pipe does not see them --
// and the per-pixel gathers become shared-memory loads (bank conflicts instead of lines).  Tiles whose box exceeds
// TILE_BOX_MAX (around the poles of the top and bottom faces a face row maps onto a wide arc) keep the direct
// gather inside the same kernel.  `rel` holds each pixel's byte offset inside its tile's box.
constexpr int TILE_H = 16, TILE_W = 128;
constexpr int TILE_BOX_MAX = 24576;

__global__ void __launch_bounds__(256)
tile_build_kernel(const int32_t* __restrict__ lut, int rows, int cols, int w, TileBox* __restrict__ tiles, uint32_t* __restrict__ rel,
                  int* __restrict__ n_fallback)
{
    __shared__ int s_rmin, s_rmax, s_cmin, s_cmax;
    __shared__ TileBox s_box;
    const int tiles_x = cols / TILE_W;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    if (threadIdx.x == 0) { s_rmin = INT_MAX; s_rmax = -1; s_cmin = INT_MAX; s_cmax = -1; }
    __syncthreads();
    constexpr int PER_THREAD = TILE_H * TILE_W / 256;
    int idx[PER_THREAD];
#pragma unroll
    for (int k = 0; k < PER_THREAD; k++) {
        const int pl = k * 256 + threadIdx.x, ly = pl >> 7, lx = pl & 127;
        const int y = ty * TILE_H + ly;
        idx[k] = (y < rows) ? __ldg(lut + (int64_t)y * cols + tx * TILE_W + lx) : -1;
        if (idx[k] >= 0) {
            const int r = idx[k] / w, cc = idx[k] - r * w;
            atomicMin(&s_rmin, r); atomicMax(&s_rmax, r);
            atomicMin(&s_cmin, cc); atomicMax(&s_cmax, cc);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        TileBox b;
        if (s_rmax < 0) { b.src_off = 0; b.nrows = 0; b.row_bytes = 0; }   // nothing to fetch (fully masked tile)
        else {
            const int c0 = (s_cmin * 3) & ~15, c1 = (s_cmax * 3 + 3 + 15) & ~15;
            const int nrows = s_rmax - s_rmin + 1, row_bytes = c1 - c0;
            if ((int64_t)nrows * row_bytes > TILE_BOX_MAX || nrows > 65535) { b.src_off = -1; b.nrows = 0; b.row_bytes = 0; atomicAdd(n_fallback, 1); }
            else { b.src_off = s_rmin * w * 3 + c0; b.nrows = (uint16_t)nrows; b.row_bytes = (uint16_t)row_bytes; }
        }
        s_box = b;
        tiles[blockIdx.x] = b;
    }
    __syncthreads();
    const TileBox b = s_box;
    const int r0 = (b.src_off >= 0 && b.nrows) ? b.src_off / (w * 3) : 0;
    const int c0b = (b.src_off >= 0 && b.nrows) ? b.src_off - r0 * w * 3 : 0;
#pragma unroll
    for (int k = 0; k < PER_THREAD; k++) {
        const int pl = k * 256 + threadIdx.x, ly = pl >> 7, lx = pl & 127;
        const int y = ty * TILE_H + ly;
        if (y >= rows) continue;
        uint32_t v = 0xFFFFFFFFu;
        if (idx[k] >= 0 && b.src_off >= 0) {
            const int r = idx[k] / w, cc = idx[k] - r * w;
            v = (uint32_t)((r - r0) * (int)b.row_bytes + (cc * 3 - c0b));
        }
        rel[(int64_t)y * cols + tx * TILE_W + lx] = v;
    }
}

template <bool MASKED>
__global__ void __launch_bounds__(256)
remap_gather_tiled_kernel(const uint8_t* __restrict__ erp, const int32_t* __restrict__ lut, const uint32_t* __restrict__ rel,
                          const TileBox* __restrict__ tiles, uint8_t* __restrict__ out, int64_t src_bytes_per_image, int src_row_bytes, int rows,
                          int cols, int tiles_x, int tiles_per_image, int n_images)
{
    __shared__ __align__(128) uint8_t box[TILE_BOX_MAX + 16];
    __shared__ __align__(16) uint32_t pix[8][128];
    __shared__ __align__(16) uint32_t stage[8][96];
    __shared__ __align__(8) uint64_t bar;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    if (threadIdx.x == 0) { bar_init(&bar, 1); bar_init_fence(); }
    __syncthreads();
    uint32_t phase = 0;
    const int64_t P = (int64_t)rows * cols;
    const int64_t last_word = src_bytes_per_image / 4 - 1;
    const int total = tiles_per_image * n_images;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        const int img = item / tiles_per_image, tile = item - img * tiles_per_image;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const TileBox tb = tiles[tile];
        const uint8_t* src = erp + (int64_t)img * src_bytes_per_image;
        const bool staged = tb.src_off >= 0;
        if (staged && tb.nrows && wib == 0) {
            if (lane == 0) bar_expect(&bar, (uint32_t)tb.nrows * tb.row_bytes);
            __syncwarp();
            for (int r = lane; r < tb.nrows; r += 32)
                bulk_load(box + (size_t)r * tb.row_bytes, src + tb.src_off + (int64_t)r * src_row_bytes, tb.row_bytes, &bar);
        }
        bool waited = !(staged && tb.nrows);
#pragma unroll 1
        for (int rr = 0; rr < TILE_H / 8; rr++) {
            const int y = ty * TILE_H + wib * (TILE_H / 8) + rr;
            if (y < rows) {
                const int64_t px0 = (int64_t)y * cols + tx * TILE_W;
                uint32_t v[4];
                if (staged) {
                    uint32_t off[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) off[k] = __ldg(rel + px0 + 32 * k + lane);
                    if (!waited) { bar_wait(&bar, phase); waited = true; }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t o = (off[k] == 0xFFFFFFFFu) ? 0u : off[k];
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(box + (o & ~3u));
                        const uint32_t sh = (o & 3u) * 8u;
                        const uint32_t w0 = wp[0], w1 = (sh > 8u) ? wp[1] : 0u;
                        v[k] = __funnelshift_r(w0, w1, sh) & 0x00FFFFFFu;
                        if (MASKED && off[k] == 0xFFFFFFFFu) v[k] = 0u;
                    }
                } else {
                    const uint32_t* words = reinterpret_cast<const uint32_t*>(src);
                    int idx[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) idx[k] = __ldg(lut + px0 + 32 * k + lane);
                    uint32_t w0[4], w1[4], sh[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int64_t a = (int64_t)((MASKED && idx[k] < 0) ? 0 : idx[k]) * 3;
                        const int64_t wi = a >> 2;
                        sh[k] = (uint32_t)(a & 3) * 8;
                        w0[k] = __ldg(words + wi);
                        w1[k] = (sh[k] > 8) ? __ldg(words + min(wi + 1, last_word)) : 0u;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        v[k] = __funnelshift_r(w0[k], w1[k], sh[k]) & 0x00FFFFFFu;
                        if (MASKED && idx[k] < 0) v[k] = 0u;
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; k++) pix[wib][32 * k + lane] = v[k];
                __syncwarp();
                const uint4 q = *reinterpret_cast<const uint4*>(&pix[wib][4 * lane]);
                stage[wib][lane * 3 + 0] = q.x | (q.y << 24);
                stage[wib][lane * 3 + 1] = (q.y >> 8) | (q.z << 16);
                stage[wib][lane * 3 + 2] = (q.z >> 16) | (q.w << 8);
                __syncwarp();
                if (lane < 24) {
                    const uint4 o = *reinterpret_cast<const uint4*>(&stage[wib][lane * 4]);
                    *reinterpret_cast<uint4*>(out + ((int64_t)img * P + px0) * 3 + lane * 16) = o;
                }
                __syncwarp();
            }
        }
        if (!waited) bar_wait(&bar, phase);      // warps whose rows fall outside the image still consume the phase
        if (staged && tb.nrows) phase ^= 1u;
        __syncthreads();                          // the box is refilled for the next tile
    }
}

// ---- source-ordered variant ---------------------------------------------------------------------------------
// What the direct gather pays for is LINES: the 32 pixels of one load instruction lie on a short output-row segment whose
// sources cross ~6 source rows on the lateral faces and ~18 near the poles (9.9 distinct 128-byte lines per load on
// average at 4K -> cube 960, and 7.5 more for the straddle load), and every line is a pass through the L1 data pipe.
// The ORDER in which a tile's pixels are fetched is free, so the plan lists the pixels of each 32 x 64 output tile sorted
// by SOURCE index: 32 consecutive entries are then a run along one or two source rows (2.5 + 2.4 lines per pixel group
// instead of 9.9 + 7.5).  An entry is (source pixel - tile base) << 11 | position in the tile, still 4 bytes per pixel.
// Pixels are scattered into a packed shared-memory image of the tile (byte stores), which leaves as aligned 16-byte
// row chunks.  Entries with all offset bits set have no source pixel (masked tables, rows past the image) and write 0.
constexpr int SORT_TH = 32, SORT_TW = 64, SORT_NPX = SORT_TH * SORT_TW, SORT_POS_BITS = 11;
constexpr uint32_t SORT_REL_NONE = 0x1FFFFFu;
constexpr int SORT_ROW_B = SORT_TW * 3;            // bytes per tile row
constexpr int SORT_CHUNKS = SORT_ROW_B / 16;        // 16-byte chunks per tile row
#ifndef SORT_BATCH
#define SORT_BATCH 4
#endif
#ifndef SORT_CTAS_PER_SM
#define SORT_CTAS_PER_SM 8
#endif

__global__ void __launch_bounds__(1024)
sorted_build_kernel(const int32_t* __restrict__ lut, int rows, int cols, int32_t* __restrict__ bases, uint32_t* __restrict__ entries,
                    int* __restrict__ n_fallback)
{
    __shared__ uint32_t key[SORT_NPX];
    __shared__ int s_min, s_max;
    const int tiles_x = cols / SORT_TW;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    if (threadIdx.x == 0) { s_min = INT_MAX; s_max = -1; }
    __syncthreads();
    int idx[2];
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int pl = k * 1024 + threadIdx.x, ly = pl / SORT_TW, lx = pl - ly * SORT_TW;
        const int y = ty * SORT_TH + ly;
        idx[k] = (y < rows) ? __ldg(lut + (int64_t)y * cols + tx * SORT_TW + lx) : -1;
        if (idx[k] >= 0) { atomicMin(&s_min, idx[k]); atomicMax(&s_max, idx[k]); }
    }
    __syncthreads();
    const int base = (s_max < 0) ? 0 : s_min;
    const bool fits = (s_max < 0) || (uint32_t)(s_max - base) < SORT_REL_NONE;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int pl = k * 1024 + threadIdx.x;
        const uint32_t rel = (idx[k] >= 0 && fits) ? (uint32_t)(idx[k] - base) : SORT_REL_NONE;
        key[pl] = (rel << SORT_POS_BITS) | (uint32_t)pl;
    }
    __syncthreads();
    // bitonic sort of the 2048 keys (position bits make them distinct)
    for (int k = 2; k <= SORT_NPX; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int t = threadIdx.x;
            const int i = 2 * t - (t & (j - 1));       // the lower index of the pair this thread owns
            const int p = i + j;
            const uint32_t a = key[i], b = key[p];
            const bool up = (i & k) == 0;
            if ((a > b) == up) { key[i] = b; key[p] = a; }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        bases[blockIdx.x] = fits ? base : -1;
        if (!fits) atomicAdd(n_fallback, 1);
    }
#pragma unroll
    for (int k = 0; k < 2; k++) entries[(int64_t)blockIdx.x * SORT_NPX + k * 1024 + threadIdx.x] = key[k * 1024 + threadIdx.x];
}

// (A variant that scattered one 32-bit word per pixel and packed rows afterwards, as the direct kernel does, halved the
// shared-memory wavefronts but measured slower: 0.256 / 0.229 ms against 0.204 ms for 16 frames.)
__device__ __forceinline__ void sorted_put(uint8_t* tb, uint32_t pos, uint32_t w0, uint32_t w1, uint32_t sh, bool none)
{
    uint32_t v = __funnelshift_r(w0, w1, sh);
    if (none) v = 0u;
    uint8_t* d = tb + pos * 3;
    d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16);
}

__global__ void __launch_bounds__(256, SORT_CTAS_PER_SM)
remap_gather_sorted_kernel(const uint8_t* __restrict__ erp, const int32_t* __restrict__ lut, const uint32_t* __restrict__ entries,
                           const int32_t* __restrict__ bases, uint8_t* __restrict__ out, int64_t src_bytes_per_image, int rows, int cols,
                           int tiles_x, int tiles_per_image, int n_images, const uint8_t* __restrict__ erp2 = nullptr,
                           uint8_t* __restrict__ out2 = nullptr, int n_first = 0)
{
    __shared__ __align__(16) uint8_t tile[2][SORT_TH * SORT_ROW_B];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int total = tiles_per_image * n_images;
    const int last_word = (int)(src_bytes_per_image / 4 - 1);
    const int64_t P = (int64_t)rows * cols;
    int buf = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        int img = item / tiles_per_image;
        const int t = item - img * tiles_per_image;
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const bool second = erp2 != nullptr && img >= n_first;
        if (second) img -= n_first;
        const uint32_t* words = reinterpret_cast<const uint32_t*>((second ? erp2 : erp) + (int64_t)img * src_bytes_per_image);
        const int base = __ldg(bases + t);
        uint8_t* tb = tile[buf];
        if (base >= 0) {
            const uint32_t* ent = entries + (int64_t)t * SORT_NPX + wib * (SORT_NPX / 8) + lane;
#pragma unroll 1
            for (int part = 0; part < 8 / SORT_BATCH; part++) {      // SORT_BATCH rounds of 32 pixels with all their loads in flight together
                uint32_t e[SORT_BATCH], w0[SORT_BATCH], w1[SORT_BATCH];
#pragma unroll
                for (int k = 0; k < SORT_BATCH; k++) e[k] = __ldg(ent + (part * SORT_BATCH + k) * 32);
#pragma unroll
                for (int k = 0; k < SORT_BATCH; k++) {
                    const uint32_t rel = e[k] >> SORT_POS_BITS;
                    const int a = (rel == SORT_REL_NONE ? 0 : base + (int)rel) * 3;     // < 2^31: checked when the plan is built
                    const int wi = a >> 2;
                    w0[k] = __ldg(words + wi);
                    w1[k] = (a & 2) ? __ldg(words + min(wi + 1, last_word)) : 0u;       // byte offset 2 or 3: the pixel straddles two words
                }
#pragma unroll
                for (int k = 0; k < SORT_BATCH; k++) {
                    const uint32_t rel = e[k] >> SORT_POS_BITS;
                    const uint32_t sh = ((rel == SORT_REL_NONE ? 0u : (uint32_t)(base + (int)rel) * 3u) & 3u) * 8u;
                    sorted_put(tb, e[k] & (SORT_NPX - 1), w0[k], w1[k], sh, rel == SORT_REL_NONE);
                }
            }
        } else {   // offsets do not fit the entry format: this tile reads the ordinary table in output order
#pragma unroll 1
            for (int k = 0; k < 8; k++) {
                const uint32_t pos = (uint32_t)(wib * (SORT_NPX / 8) + k * 32 + lane);
                const int ly = pos / SORT_TW, lx = pos - ly * SORT_TW;
                const int y = ty * SORT_TH + ly;
                const int src = (y < rows) ? __ldg(lut + (int64_t)y * cols + tx * SORT_TW + lx) : -1;
                const int a = (src < 0 ? 0 : src) * 3;
                const int wi = a >> 2;
                const uint32_t w0 = __ldg(words + wi);
                const uint32_t w1 = (a & 2) ? __ldg(words + min(wi + 1, last_word)) : 0u;
                sorted_put(tb, pos, w0, w1, (uint32_t)(a & 3) * 8u, src < 0);
            }
        }
        __syncthreads();
        uint8_t* dst0 = (second ? out2 : out) + ((int64_t)img * P + (int64_t)ty * SORT_TH * cols + (int64_t)tx * SORT_TW) * 3;
        for (int ch = threadIdx.x; ch < SORT_TH * SORT_CHUNKS; ch += 256) {
            const int r = ch / SORT_CHUNKS, cc = ch - r * SORT_CHUNKS;
            if (ty * SORT_TH + r < rows)
                *reinterpret_cast<uint4*>(dst0 + (int64_t)r * cols * 3 + cc * 16) = *reinterpret_cast<const uint4*>(tb + r * SORT_ROW_B + cc * 16);
        }
        buf ^= 1;   // the next tile is scattered into the other buffer; the barrier above orders the one after that
    }
}

void free_tiled_plan(TiledPlan* tp)
{
    if (tp->tiles) cudaFree(tp->tiles);
    if (tp->rel) cudaFree(tp->rel);
    if (tp->sorted) cudaFree(tp->sorted);
    if (tp->sorted_base) cudaFree(tp->sorted_base);
    *tp = TiledPlan();
}

int build_tiled_plan(sba_ctx* c, const int32_t* lut, int rows, int cols, int w, int h, bool masked, TiledPlan* out)
{
    *out = TiledPlan();
    if ((int64_t)w * h * 3 >= ((int64_t)1 << 31)) return SBA_OK;   // both forms keep source byte offsets in 31 bits
    TiledPlan tp;
    SBA_TRY(c->scratch[SCR_WORK1].ensure(2 * sizeof(int), c->stream));
    int* d_cnt = c->scratch[SCR_WORK1].as<int>();
    SBA_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), c->stream));
    // tiled form: whole 128-pixel tile columns, 16-byte aligned source rows
    if (cols % TILE_W == 0 && (w * 3) % 16 == 0) {
        tp.tiles_x = cols / TILE_W;
        tp.tiles_y = (rows + TILE_H - 1) / TILE_H;
        const int n_tiles = tp.tiles_x * tp.tiles_y;
        SBA_CUDA(cudaMalloc(&tp.tiles, (size_t)n_tiles * sizeof(TileBox)));
        SBA_CUDA(cudaMalloc(&tp.rel, (size_t)rows * cols * sizeof(uint32_t)));
        tile_build_kernel<<<n_tiles, 256, 0, c->stream>>>(lut, rows, cols, w, tp.tiles, tp.rel, d_cnt);
        SBA_LAUNCHED(c);
    }
    // source-ordered form: whole 64-pixel tile columns
    if (cols % SORT_TW == 0) {
        tp.s_tiles_x = cols / SORT_TW;
        tp.s_tiles_y = (rows + SORT_TH - 1) / SORT_TH;
        const int n_tiles = tp.s_tiles_x * tp.s_tiles_y;
        SBA_CUDA(cudaMalloc(&tp.sorted, (size_t)n_tiles * SORT_NPX * sizeof(uint32_t)));
        SBA_CUDA(cudaMalloc(&tp.sorted_base, (size_t)n_tiles * sizeof(int32_t)));
        sorted_build_kernel<<<n_tiles, 1024, 0, c->stream>>>(lut, rows, cols, tp.sorted_base, tp.sorted, d_cnt + 1);
        SBA_LAUNCHED(c);
    }
    SBA_CUDA(cudaGetLastError());
    int cnt[2];
    SBA_CUDA(cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, c->stream));
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    tp.n_fallback = cnt[0];
    tp.s_fallback = cnt[1];
    if (!tp.tiles && !tp.sorted) { *out = tp; return SBA_OK; }

    // Which kernel is fastest depends on the table (how many source rows a tile's box spans decides how many bulk
    // copies the tiled kernel takes, how long the source runs of a tile are decides what source order buys) and on
    // whether the sources stream from HBM or sit in L2, so the kernels are timed once per plan on scratch frames:
    // 2 frames (the per-pair calls) and a batch larger than L2.
    const size_t src_bytes = (size_t)w * h * 3, out_bytes = (size_t)rows * cols * 3;
    const int big = (int)std::min<size_t>(8, std::max<size_t>(3, ((size_t)160 << 20) / src_bytes + 1));
    SBA_TRY(c->scratch[SCR_WORK3].ensure(big * src_bytes, c->stream));
    SBA_TRY(c->scratch[SCR_WORK4].ensure(big * out_bytes, c->stream));
    SBA_CUDA(cudaMemsetAsync(c->scratch[SCR_WORK3].p, 0x5a, big * src_bytes, c->stream));
    cudaEvent_t e0, e1;
    SBA_CUDA(cudaEventCreate(&e0));
    SBA_CUDA(cudaEventCreate(&e1));
    const int saved_mode = c->remap_kernel;
    int rc = SBA_OK;
    for (int cls = 0; cls < 2 && rc == SBA_OK; cls++) {
        const int frames = cls == 0 ? 2 : big;
        for (int variant = 0; variant < 3 && rc == SBA_OK; variant++) {
            if ((variant == 1 && !tp.tiles) || (variant == 2 && !tp.sorted)) { tp.trial_ms[cls][variant] = 0.f; continue; }
            c->remap_kernel = variant + 1;
            float best = 1e30f;
            const int reps = cls == 0 ? 9 : 4;   // the 2-frame launches are ~35 us each: more of them, the minimum is what counts
            for (int rep = 0; rep < reps && rc == SBA_OK; rep++) {
                cudaEventRecord(e0, c->stream);
                rc = launch_lut_gather(c, c->scratch[SCR_WORK3].as<uint8_t>(), (int64_t)src_bytes, lut, rows, cols, c->scratch[SCR_WORK4].as<uint8_t>(),
                                       frames, masked, &tp, w);
                cudaEventRecord(e1, c->stream);
                cudaEventSynchronize(e1);
                float ms = 0.f;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0) best = std::min(best, ms);
            }
            tp.trial_ms[cls][variant] = best;
        }
        // a challenger has to win by 3 % to displace the direct gather
        tp.best[cls] = 1;
        float best_ms = 0.97f * tp.trial_ms[cls][0];
        for (int variant = 1; variant < 3; variant++)
            if (tp.trial_ms[cls][variant] > 0.f && tp.trial_ms[cls][variant] < best_ms) { best_ms = tp.trial_ms[cls][variant]; tp.best[cls] = variant + 1; }
        tp.preferred[cls] = tp.best[cls] == 2;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    c->remap_kernel = saved_mode;
    c->scratch[SCR_WORK3].release();   // the trial buffers are large and never needed again
    c->scratch[SCR_WORK4].release();
    if (rc != SBA_OK) { free_tiled_plan(&tp); return rc; }
    *out = tp;
    return SBA_OK;
}

// Generic path: one thread per pixel; handles odd sizes, unaligned bases and single faces
// (`lut_row_stride`/`face_off` select a cs-wide window of the strip table).
__global__ void remap_gather1_kernel(const uint8_t* __restrict__ erp, const int32_t* __restrict__ lut, uint8_t* __restrict__ out,
                                     int64_t src_bytes_per_image, int rows, int cols, int lut_row_stride, int face_off, int n_images,
                                     int64_t p_first)
{
    // pixels [p_first, P) of every image
    int64_t P = (int64_t)rows * cols, per = P - p_first, total = per * n_images;
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        int img = (int)(g / per);
        int64_t p = p_first + (g - (int64_t)img * per);
        int i = (int)(p / cols), j = (int)(p - (int64_t)i * cols);
        int32_t s = __ldg(lut + (int64_t)i * lut_row_stride + face_off + j);
        uint8_t* o = out + ((int64_t)img * P + p) * 3;
        if (s < 0) { o[0] = o[1] = o[2] = 0; continue; }   // masked tables only (the cube tables hold no negative entry)
        const uint8_t* q = erp + (int64_t)img * src_bytes_per_image + (int64_t)s * 3;
        o[0] = __ldg(q); o[1] = __ldg(q + 1); o[2] = __ldg(q + 2);
    }
}

// ---- elementwise keypoint kernels --------------------------------------------------------------
// equi2cube_surf.cpp:19-76
__global__ void cube2equi_points_kernel(const float2* __restrict__ in, int n, int cs, int w, int h, float2* __restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float2 p = in[k];
    float2 o;
    cube2equi_point(p.x, p.y, cs, w, h, &o.x, &o.y);
    out[k] = o;
}

// spherical_bundle_adjuster.cpp:271-298
__global__ void pixels_to_bearings_kernel(const float2* __restrict__ px, int n, double w, double h, float4* __restrict__ b32,
                                          double* __restrict__ b64)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float2 p = px[k];
    double x, y, z;
    pixel_to_bearing(p.x, p.y, w, h, &x, &y, &z);
    if (b32) b32[k] = make_float4((float)x, (float)y, (float)z, 0.f);
    if (b64) { b64[3 * k] = x; b64[3 * k + 1] = y; b64[3 * k + 2] = z; }
}

// equi2cube_surf.cpp:107-113
__global__ void gather_matches_kernel(const float2* __restrict__ kl, const float2* __restrict__ kr, const int32_t* __restrict__ qi,
                                      const int32_t* __restrict__ ti, int n, float2* __restrict__ ol, float2* __restrict__ orr)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ol[k] = kl[qi[k]];
    orr[k] = kr[ti[k]];
}

// Table-driven gather of n_images images: out[img][p] = erp[img][lut[p]] for p in [0, rows*cols).
int launch_lut_gather(sba_ctx* c, const uint8_t* d_erp, int64_t src_bytes, const int32_t* lut, int rows, int cols, uint8_t* d_out, int n_images,
                      bool masked, const TiledPlan* tiled, int src_w)
{
    const int64_t P = (int64_t)rows * cols;
    const int choice = c->remap_kernel != 0 ? c->remap_kernel : (tiled ? tiled->best[n_images >= 3 ? 1 : 0] : 1);
    if (choice == 3 && tiled && tiled->sorted && (uintptr_t)d_erp % 4 == 0 && src_bytes % 4 == 0 && (uintptr_t)d_out % 16 == 0 &&
        ((P * 3) % 16 == 0 || n_images == 1) && (int64_t)tiled->s_tiles_x * tiled->s_tiles_y * n_images < ((int64_t)1 << 31)) {
        const int per_image = tiled->s_tiles_x * tiled->s_tiles_y, total = per_image * n_images;
        const int blocks = std::min(total, c->sm_count * SORT_CTAS_PER_SM);
        remap_gather_sorted_kernel<<<blocks, 256, 0, c->stream>>>(d_erp, lut, tiled->sorted, tiled->sorted_base, d_out, src_bytes, rows, cols,
                                                                  tiled->s_tiles_x, per_image, n_images);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
        return SBA_OK;
    }
    const bool want_tiled = tiled && tiled->tiles && choice == 2;
    const bool fast_ok = ((uintptr_t)d_erp % 4 == 0) && (src_bytes % 4 == 0) && ((uintptr_t)d_out % 16 == 0) && ((uintptr_t)lut % 16 == 0) &&
                         ((P * 3) % 16 == 0 || n_images == 1) && P >= 128;
    int64_t done_px = 0;
    if (fast_ok) {
        const int64_t groups = P / 128;
        const int per_launch = (int)std::max<int64_t>(1, std::min<int64_t>(n_images, ((int64_t)1 << 30) / groups));   // 32-bit group ids in the kernel
        for (int first = 0; first < n_images; first += per_launch) {
            const int nimg = std::min(per_launch, n_images - first);
            const int64_t total = groups * nimg;
            // persistent-style grid: a multiple of the SM count, 8 resident CTAs (64 warps) per SM
            int blocks = (int)std::min<int64_t>(ceil_div64(total, 8), (int64_t)c->sm_count * 8);
            if (blocks >= c->sm_count) blocks = blocks / c->sm_count * c->sm_count;
            const uint8_t* src = d_erp + (int64_t)first * src_bytes;
            uint8_t* dst = d_out + (int64_t)first * P * 3;
            if (masked) remap_gather_warp_kernel<true><<<blocks, 256, 0, c->stream>>>(src, lut, dst, src_bytes, P, (int)groups, nimg);
            else remap_gather_warp_kernel<false><<<blocks, 256, 0, c->stream>>>(src, lut, dst, src_bytes, P, (int)groups, nimg);
            SBA_LAUNCHED(c);
        }
        done_px = groups * 128;
    }
    if (done_px < P) {   // tail of every image (or everything when the fast path does not apply)
        int64_t total = (P - done_px) * n_images;
        int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)c->sm_count * 32);
        remap_gather1_kernel<<<blocks, 256, 0, c->stream>>>(d_erp, lut, d_out, src_bytes, rows, cols, cols, 0, n_images, done_px);
        SBA_LAUNCHED(c);
    }
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

// equi2cube::get_all of the two images of a pair (two separate device buffers in, two out) with one launch when the fast
// kernel applies; otherwise two ordinary calls.
int equi2cube_pair(sba_ctx* c, const uint8_t* d_im0, const uint8_t* d_im1, int w, int h, int cs, uint8_t* d_s0, uint8_t* d_s1)
{
    RemapPlan* plan;
    SBA_TRY(get_plan(c, w, h, cs, &plan));
    const int64_t P = (int64_t)cs * 6 * cs, src_bytes = (int64_t)w * h * 3;
    auto ok4 = [](const void* p) { return (uintptr_t)p % 4 == 0; };
    auto ok16 = [](const void* p) { return (uintptr_t)p % 16 == 0; };
    const int choice = c->remap_kernel != 0 ? c->remap_kernel : plan->tiled.best[0];
    if (choice == 3 && plan->tiled.sorted && ok4(d_im0) && ok4(d_im1) && src_bytes % 4 == 0 && ok16(d_s0) && ok16(d_s1)) {
        const TiledPlan& tp = plan->tiled;
        const int per_image = tp.s_tiles_x * tp.s_tiles_y;
        const int blocks = std::min(per_image * 2, c->sm_count * SORT_CTAS_PER_SM);
        prof_begin(c, SBA_KERNEL_REMAP);
        remap_gather_sorted_kernel<<<blocks, 256, 0, c->stream>>>(d_im0, plan->lut, tp.sorted, tp.sorted_base, d_s0, src_bytes, cs, 6 * cs, tp.s_tiles_x,
                                                                  per_image, 2, d_im1, d_s1, 1);
        prof_end(c, SBA_KERNEL_REMAP);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
        return SBA_OK;
    }
    if (choice != 2 && ok4(d_im0) && ok4(d_im1) && src_bytes % 4 == 0 && ok16(d_s0) && ok16(d_s1) && ok16(plan->lut) && P % 128 == 0 &&
        P / 128 < ((int64_t)1 << 29)) {
        const int groups = (int)(P / 128);
        int blocks = (int)std::min<int64_t>(ceil_div64((int64_t)groups * 2, 8), (int64_t)c->sm_count * 8);
        if (blocks >= c->sm_count) blocks = blocks / c->sm_count * c->sm_count;
        prof_begin(c, SBA_KERNEL_REMAP);
        remap_gather_warp_kernel<false><<<blocks, 256, 0, c->stream>>>(d_im0, plan->lut, d_s0, src_bytes, P, groups, 2, d_im1, d_s1, 1);
        prof_end(c, SBA_KERNEL_REMAP);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
        return SBA_OK;
    }
    SBA_TRY(sba_equi2cube(c, d_im0, w, h, 1, cs, d_s0, SBA_MEM_DEVICE));
    return sba_equi2cube(c, d_im1, w, h, 1, cs, d_s1, SBA_MEM_DEVICE);
}

static int launch_gather(sba_ctx* c, const uint8_t* d_erp, const RemapPlan* plan, uint8_t* d_out, int n_images, int face)
{
    int cs = plan->cs;
    int64_t src_bytes = (int64_t)plan->w * plan->h * 3;
    prof_begin(c, SBA_KERNEL_REMAP);
    if (face < 0) {
        SBA_TRY(launch_lut_gather(c, d_erp, src_bytes, plan->lut, cs, 6 * cs, d_out, n_images, false, &plan->tiled, plan->w));
    } else {
        int64_t total = (int64_t)cs * cs * n_images;
        int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)c->sm_count * 32);
        remap_gather1_kernel<<<blocks, 256, 0, c->stream>>>(d_erp, plan->lut, d_out, src_bytes, cs, cs, 6 * cs, face * cs, n_images, 0);
        SBA_LAUNCHED(c);
    }
    prof_end(c, SBA_KERNEL_REMAP);
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

}  // namespace sba

using namespace sba;

extern "C" {

int sba_ctx_set_remap_kernel(sba_ctx* c, int mode)
{
    SBA_CHECK_ARG(c && mode >= 0 && mode <= 3);
    c->remap_kernel = mode;
    return SBA_OK;
}

int sba_remap_plan_info(sba_ctx* c, int w, int h, int cs, int* tiled_available, int* tiled_preferred, int* n_tiles, int* n_fallback_tiles,
                        float trial_ms[4])
{
    SBA_CHECK_ARG(c && w > 0 && h > 0 && cs > 0);
    SBA_CUDA(cudaSetDevice(c->device));
    RemapPlan* plan;
    SBA_TRY(get_plan(c, w, h, cs, &plan));
    const TiledPlan& tp = plan->tiled;
    if (tiled_available) *tiled_available = tp.tiles != nullptr;
    if (tiled_preferred) *tiled_preferred = (tp.preferred[0] ? 1 : 0) | (tp.preferred[1] ? 2 : 0);
    if (n_tiles) *n_tiles = tp.tiles_x * tp.tiles_y;
    if (n_fallback_tiles) *n_fallback_tiles = tp.n_fallback;
    if (trial_ms) for (int k = 0; k < 4; k++) trial_ms[k] = tp.trial_ms[k >> 1][k & 1];
    return SBA_OK;
}

int sba_remap_plan_sorted_info(sba_ctx* c, int w, int h, int cs, int* available, int* best_small, int* best_large, int* n_tiles,
                               int* n_fallback_tiles, float trial_ms[2])
{
    SBA_CHECK_ARG(c && w > 0 && h > 0 && cs > 0);
    SBA_CUDA(cudaSetDevice(c->device));
    RemapPlan* plan;
    SBA_TRY(get_plan(c, w, h, cs, &plan));
    const TiledPlan& tp = plan->tiled;
    if (available) *available = tp.sorted != nullptr;
    if (best_small) *best_small = tp.best[0];
    if (best_large) *best_large = tp.best[1];
    if (n_tiles) *n_tiles = tp.s_tiles_x * tp.s_tiles_y;
    if (n_fallback_tiles) *n_fallback_tiles = tp.s_fallback;
    if (trial_ms) { trial_ms[0] = tp.trial_ms[0][2]; trial_ms[1] = tp.trial_ms[1][2]; }
    return SBA_OK;
}

int sba_equi2cube_lut(sba_ctx* c, int w, int h, int cs, int32_t* lut_out, int mem)
{
    SBA_CHECK_ARG(c && w > 0 && h > 0 && cs > 0);
    SBA_CUDA(cudaSetDevice(c->device));
    RemapPlan* plan;
    SBA_TRY(get_plan(c, w, h, cs, &plan));
    if (lut_out) {
        size_t bytes = (size_t)cs * 6 * cs * sizeof(int32_t);
        SBA_CUDA(cudaMemcpyAsync(lut_out, plan->lut, bytes, mem == SBA_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    }
    return finish(c, mem);
}

int sba_equi2cube(sba_ctx* c, const uint8_t* erp, int w, int h, int n_images, int cs, uint8_t* strip_out, int mem)
{
    SBA_CHECK_ARG(c && erp && strip_out && w > 0 && h > 0 && cs > 0 && n_images >= 0);
    if (n_images == 0) return SBA_OK;
    SBA_CUDA(cudaSetDevice(c->device));
    RemapPlan* plan;
    SBA_TRY(get_plan(c, w, h, cs, &plan));
    size_t in_bytes = (size_t)w * h * 3 * n_images, out_bytes = (size_t)cs * 6 * cs * 3 * n_images;
    const uint8_t* d_in;
    uint8_t* d_out;
    SBA_TRY(stage_in(c, erp, in_bytes, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, strip_out, out_bytes, mem, SCR_OUT0, &d_out));
    SBA_TRY(launch_gather(c, d_in, plan, d_out, n_images, -1));
    SBA_TRY(copy_out(c, strip_out, d_out, out_bytes, mem));
    return finish(c, mem);
}

int sba_equi2cube_face(sba_ctx* c, const uint8_t* erp, int w, int h, int cs, int face, uint8_t* face_out, int mem)
{
    SBA_CHECK_ARG(c && erp && face_out && w > 0 && h > 0 && cs > 0 && face >= 0 && face < 6);
    SBA_CUDA(cudaSetDevice(c->device));
    RemapPlan* plan;
    SBA_TRY(get_plan(c, w, h, cs, &plan));
    size_t in_bytes = (size_t)w * h * 3, out_bytes = (size_t)cs * cs * 3;
    const uint8_t* d_in;
    uint8_t* d_out;
    SBA_TRY(stage_in(c, erp, in_bytes, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, face_out, out_bytes, mem, SCR_OUT0, &d_out));
    SBA_TRY(launch_gather(c, d_in, plan, d_out, 1, face));
    SBA_TRY(copy_out(c, face_out, d_out, out_bytes, mem));
    return finish(c, mem);
}

int sba_cube2equi_points(sba_ctx* c, const float* xy_in, int n, int cs, int w, int h, float* xy_out, int mem)
{
    SBA_CHECK_ARG(c && n >= 0 && cs > 0 && w > 0 && h > 0);
    if (n == 0) return SBA_OK;
    SBA_CHECK_ARG(xy_in && xy_out);
    SBA_CUDA(cudaSetDevice(c->device));
    const float* d_in;
    float* d_out;
    SBA_TRY(stage_in(c, xy_in, (size_t)2 * n, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, xy_out, (size_t)2 * n, mem, SCR_OUT0, &d_out));
    cube2equi_points_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const float2*)d_in, n, cs, w, h, (float2*)d_out);
    SBA_LAUNCHED(c);
    SBA_TRY(copy_out(c, xy_out, d_out, (size_t)2 * n, mem));
    return finish(c, mem);
}

int sba_pixels_to_bearings(sba_ctx* c, const float* xy, int n, int w, int h, float* b32, double* b64, int mem)
{
    SBA_CHECK_ARG(c && n >= 0 && w > 0 && h > 0);
    if (n == 0) return SBA_OK;
    SBA_CHECK_ARG(xy && (b32 || b64));
    SBA_CUDA(cudaSetDevice(c->device));
    const float* d_in;
    float* d_b32;
    double* d_b64;
    SBA_TRY(stage_in(c, xy, (size_t)2 * n, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, b32, (size_t)4 * n, mem, SCR_OUT0, &d_b32));
    SBA_TRY(stage_out(c, b64, (size_t)3 * n, mem, SCR_OUT1, &d_b64));
    pixels_to_bearings_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const float2*)d_in, n, (double)w, (double)h, (float4*)d_b32, d_b64);
    SBA_LAUNCHED(c);
    SBA_TRY(copy_out(c, b32, d_b32, (size_t)4 * n, mem));
    SBA_TRY(copy_out(c, b64, d_b64, (size_t)3 * n, mem));
    return finish(c, mem);
}

int sba_gather_matches(sba_ctx* c, const float* kl, const float* kr, const int32_t* qi, const int32_t* ti, int n, float* ol, float* orr, int mem)
{
    SBA_CHECK_ARG(c && n >= 0);
    if (n == 0) return SBA_OK;
    SBA_CHECK_ARG(kl && kr && qi && ti && ol && orr);
    if (mem != SBA_MEM_DEVICE) {
        // The keypoint array lengths are not part of this call, so host arrays cannot be staged;
        // the facade gathers its own cv::KeyPoint vectors, the device pipeline calls this.
        set_error("sba_gather_matches takes device pointers only");
        return SBA_ERR_UNSUPPORTED;
    }
    SBA_CUDA(cudaSetDevice(c->device));
    gather_matches_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const float2*)kl, (const float2*)kr, qi, ti, n, (float2*)ol, (float2*)orr);
    SBA_LAUNCHED(c);
    return finish(c, mem);
}

}  // extern "C"
