// common.cuh -- context, error handling and scratch memory shared by all kernels of libsba_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <utility>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "sba_b200.h"

namespace sba {

void set_error(const char* fmt, ...);

#define SBA_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            sba::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return SBA_ERR_CUDA;                                                                   \
        }                                                                                          \
    } while (0)

#define SBA_CHECK_ARG(cond)                                                         \
    do {                                                                            \
        if (!(cond)) {                                                              \
            sba::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
            return SBA_ERR_INVALID;                                                 \
        }                                                                           \
    } while (0)

#define SBA_TRY(expr)            \
    do {                         \
        int s__ = (expr);        \
        if (s__ != SBA_OK) return s__; \
    } while (0)

// Grow-only device buffer.  Growing synchronises the stream first (the old block may be in use).
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes, cudaStream_t s)
    {
        if (bytes <= cap) return SBA_OK;
        if (p) {
            SBA_CUDA(cudaStreamSynchronize(s));
            SBA_CUDA(cudaFree(p));
            p = nullptr; cap = 0;
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            p = nullptr;
            set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
            return SBA_ERR_NOMEM;
        }
        cap = want;
        return SBA_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
    }
    template <typename T> T* as() { return (T*)p; }
};

// Size-bucketed cache of device / pinned-host blocks.  Everything a context allocates per call (BA
// problems are created and destroyed once per image pair) is recycled here, so steady-state calls do
// no cudaMalloc.  All users enqueue on the context's single stream, so reuse is stream-ordered.
struct BlockCache {
    std::multimap<size_t, void*> free_dev, free_host;
    std::map<void*, size_t> live_dev, live_host;
    static size_t bucket(size_t bytes)
    {
        size_t b = 256;
        while (b < bytes) b <<= 1;
        return b;
    }
    cudaError_t get(void** p, size_t bytes, bool host)
    {
        const size_t b = bucket(bytes ? bytes : 1);
        auto& fl = host ? free_host : free_dev;
        auto it = fl.find(b);
        if (it != fl.end()) {
            *p = it->second;
            fl.erase(it);
        } else {
            cudaError_t e = host ? cudaMallocHost(p, b) : cudaMalloc(p, b);
            if (e != cudaSuccess) { *p = nullptr; return e; }
        }
        (host ? live_host : live_dev)[*p] = b;
        return cudaSuccess;
    }
    void put(void* p, bool host)
    {
        if (!p) return;
        auto& live = host ? live_host : live_dev;
        auto it = live.find(p);
        if (it == live.end()) return;
        (host ? free_host : free_dev).emplace(it->second, p);
        live.erase(it);
    }
    void release_all()
    {
        for (auto& kv : free_dev) cudaFree(kv.second);
        for (auto& kv : live_dev) cudaFree(kv.first);
        for (auto& kv : free_host) cudaFreeHost(kv.second);
        for (auto& kv : live_host) cudaFreeHost(kv.first);
        free_dev.clear(); free_host.clear(); live_dev.clear(); live_host.clear();
    }
};

// Second form of a gather table for the tiled kernel (remap.cu): per 16 x 128 output tile the bounding box of its
// source pixels (staged in shared memory by bulk copies) and per pixel its byte offset inside that box.
struct TileBox {
    int32_t src_off;      // byte offset of the box inside one source image (16-byte aligned); < 0: box too large, gather directly
    uint16_t nrows;       // source rows covered
    uint16_t row_bytes;   // bytes staged per row (multiple of 16)
};
struct TiledPlan {
    TileBox* tiles = nullptr;   // device, tiles_y x tiles_x
    uint32_t* rel = nullptr;    // device, rows x cols; 0xFFFFFFFF = no source pixel
    int tiles_x = 0, tiles_y = 0;
    int n_fallback = 0;
    // set by timed trials at plan construction: does the tiled kernel beat the direct gather for this table -- on a couple
    // of frames (sources stay in L2) and on a batch that streams from HBM (8 frames)?  [0] = small, [1] = large batches
    bool preferred[2] = {false, false};
    float trial_ms[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // [batch class][direct, tiled, source-ordered]
    // Third form (remap_gather_sorted_kernel): per 32 x 64 output tile its pixels listed in SOURCE order, one 32-bit entry
    // each = (source pixel - tile base) << 11 | position inside the tile.
    uint32_t* sorted = nullptr;      // device, s_tiles x 2048 entries
    int32_t* sorted_base = nullptr;  // device, per tile: smallest source pixel index; < 0: offsets do not fit, gather from the table
    int s_tiles_x = 0, s_tiles_y = 0;
    int s_fallback = 0;
    int best[2] = {1, 1};            // fastest kernel per batch class from the trials: 1 direct, 2 tiled, 3 source-ordered
};

struct RemapPlan {
    int w, h, cs;
    TiledPlan tiled;
    int32_t* lut = nullptr;  // device, cs x 6cs
    int n_patched = 0;       // near-integer pixels resolved with the host libm
    int n_clamped = 0;
};

// spherical_surf front-end: source index of every pixel of the cropped band for one pitch (spherical.cu)
struct CropPlan {
    int w, h;
    float pitch_deg;
    int32_t* lut = nullptr;   // device, (h/4) x w; -1 = the reference's bounds check fails
    int n_patched = 0;
};

// Named scratch slots (one DevBuf each) so independent stages never alias.
enum ScratchSlot {
    SCR_IN0 = 0, SCR_IN1, SCR_IN2, SCR_IN3, SCR_OUT0, SCR_OUT1, SCR_OUT2, SCR_OUT3, SCR_OUT4,
    SCR_WORK0, SCR_WORK1, SCR_WORK2, SCR_WORK3, SCR_WORK4, SCR_WORK5,
    // buffers owned by the fused pair pipeline (pipeline.cu); the stage entry points never touch them
    SCR_PIPE_IM0, SCR_PIPE_IM1, SCR_PIPE_STRIP0, SCR_PIPE_STRIP1, SCR_PIPE_DESC0, SCR_PIPE_DESC1, SCR_PIPE_KEY0, SCR_PIPE_KEY1,
    SCR_PIPE_MATCH, SCR_PIPE_PTS, SCR_PIPE_BEAR,
    SCR_FIN_COUNTS,   // epoch-tagged survivor counts of knn2_finalize_kernel
    SCR_COUNT
};

}  // namespace sba

struct sba_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    int64_t launches = 0;
    sba::DevBuf scratch[sba::SCR_COUNT];
    sba::BlockCache cache;
    std::map<std::tuple<int, int, int>, sba::RemapPlan> plans;
    std::map<std::tuple<int, int, uint32_t>, sba::CropPlan> crop_plans;   // keyed by (w, h, bits of the pitch)
    std::map<std::pair<int, int>, int32_t*> band_plans;                  // the four bands of spherical_surf::do_all in one table
    std::map<std::pair<int, int>, sba::TiledPlan> band_tiled;             // ... and its tiled form
    std::map<std::tuple<int, int, int, int>, std::pair<std::vector<int>, int*>> tc_spans;   // tensor matcher: per-geometry query-block -> CTA span tables (matcher_tc.cu)
    sba_match_stats match_stats{};
    unsigned int fin_epoch = 0;   // call counter of knn2_finalize_kernel (tags its count slots)
    const void* fb_parts = nullptr;   // tensor matcher -> finalize: partial top-2 lists of the exact fallback (Top2*), queue length, scan grid
    const int* fb_count = nullptr;
    int fb_grid = 0;
    int matcher_ctas = 0;       // persistent CTAs of the tensor-core matcher; 0 = one per SM (sba_ctx_set_matcher_ctas)
    bool pdl_small = true;      // dependent launch for the SMALL kernels of the chain (fallback, finalize, pair solve): their early CTAs hog nothing, so it is on by default
    bool pdl = false;           // programmatic dependent launch along the pair chain (sba_ctx_set_dependent_launch)
    int remap_kernel = 0;       // 0 = per-plan choice from the timed trial, 1 = direct gather, 2 = tiled, 3 = source-ordered (sba_ctx_set_remap_kernel)
    bool pair_pending = false;  // a sba_pair_rotation_begin whose _end has not run yet (pipeline.cu)
    int* pinned_i32 = nullptr;  // small pinned host mailbox (64 ints) for scalar read-backs
    bool profiling = false;
    cudaEvent_t prof_e0[3] = {nullptr, nullptr, nullptr}, prof_e1[3] = {nullptr, nullptr, nullptr};
    bool prof_valid[3] = {false, false, false};
    // second stream + events: bulk host->device image uploads overlap compute in sba_pair_rotation
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_ev[2] = {nullptr, nullptr};
    cudaEvent_t main_ev = nullptr;
};

namespace sba {

// Where a user pointer lives decides whether we stage it.  Returns the device pointer to use.
template <typename T>
inline int stage_in(sba_ctx* c, const T* user, size_t count, int mem, ScratchSlot slot, const T** dev)
{
    if (count == 0 || user == nullptr) { *dev = user; return SBA_OK; }
    if (mem == SBA_MEM_DEVICE) { *dev = user; return SBA_OK; }
    SBA_TRY(c->scratch[slot].ensure(count * sizeof(T), c->stream));
    SBA_CUDA(cudaMemcpyAsync(c->scratch[slot].p, user, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *dev = (const T*)c->scratch[slot].p;
    return SBA_OK;
}

template <typename T>
inline int stage_out(sba_ctx* c, T* user, size_t count, int mem, ScratchSlot slot, T** dev)
{
    if (user == nullptr) { *dev = nullptr; return SBA_OK; }
    if (mem == SBA_MEM_DEVICE) { *dev = user; return SBA_OK; }
    SBA_TRY(c->scratch[slot].ensure((count ? count : 1) * sizeof(T), c->stream));
    *dev = (T*)c->scratch[slot].p;
    return SBA_OK;
}

template <typename T>
inline int copy_out(sba_ctx* c, T* user, const T* dev, size_t count, int mem)
{
    if (user == nullptr || mem == SBA_MEM_DEVICE || count == 0) return SBA_OK;
    SBA_CUDA(cudaMemcpyAsync(user, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    return SBA_OK;
}

inline int finish(sba_ctx* c, int mem)
{
    SBA_CUDA(cudaGetLastError());
    if (mem == SBA_MEM_HOST) SBA_CUDA(cudaStreamSynchronize(c->stream));
    return SBA_OK;
}

#define SBA_LAUNCHED(ctx) ((ctx)->launches++)

// ---- programmatic dependent launch (the kernels of one pair are a chain on one stream) ----------------------------------
// A kernel launched with launch_pdl may be scheduled while its predecessor in the stream is still draining (its CTAs take
// SMs as the predecessor's retire, so the ~4 us between two dependent launches disappears).  Such a kernel calls pdl_wait()
// before it reads or writes ANY global memory: the call returns when the predecessor grid has completed and its writes are
// visible -- and since every kernel of the chain waits like that, when all earlier ones have.  pdl_trigger() at the top of a
// kernel lets the NEXT launch start being scheduled; without the launch attribute both calls are no-ops.
// Measured on B200 (C2 pair): one pair at a time 254 -> 240 us, but with six pairs in flight 7 706 -> 7 506 pairs/s -- CTAs
// that sit on an SM waiting for their own chain keep other pairs' kernels off it -- so it is a per-context latency knob
// (sba_ctx_set_dependent_launch, default off).  The three small kernels at the end of the chain (fallback, finalize, pair solve)
// are launched that way always: 0.233 -> 0.223 ms one pair at a time, throughput unchanged (8 257 / 8 160 vs 8 271 / 8 194 pairs/s).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(bool enable, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    const bool off = !enable;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = off ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#endif

// Bracket the dominant kernel of a stage with events when profiling is on.
inline void prof_begin(sba_ctx* c, int id)
{
    if (c->profiling) cudaEventRecord(c->prof_e0[id], c->stream);
}
inline void prof_end(sba_ctx* c, int id)
{
    if (c->profiling) {
        cudaEventRecord(c->prof_e1[id], c->stream);
        c->prof_valid[id] = true;
    }
}

// ba.cu: problem creation with the option to read the caller's device bearings in place
int ba_problem_create_impl(sba_ctx* c, const float* b1, const float* b2, const int32_t* cam, int64_t n_obs, int n_cam, int mem, bool borrow,
                           const int* d_n_obs, sba_ba_problem** out);

// ba.cu: what the other translation units may see of a problem (ba_depth.cu runs its own kernels over the
// same observations)
struct BaView {
    sba_ctx* ctx;
    const float4* b1;
    const float4* b2;
    int64_t n_obs;
    int n_cam;
    const int* n_obs_dev;
};
BaView ba_problem_view(sba_ba_problem* p);

// remap.cu: out[img][p] = erp[img][lut[p]] (3-byte pixels); masked tables hold -1 for "no source", which gives 0
int launch_lut_gather(sba_ctx* c, const uint8_t* d_erp, int64_t src_bytes, const int32_t* lut, int rows, int cols, uint8_t* d_out, int n_images,
                      bool masked, const TiledPlan* tiled = nullptr, int src_w = 0);
// remap.cu: both images of a pair (separate device buffers) through equi2cube::get_all, one launch when possible
int equi2cube_pair(sba_ctx* c, const uint8_t* d_im0, const uint8_t* d_im1, int w, int h, int cs, uint8_t* d_s0, uint8_t* d_s1);
// remap.cu: derive the tiled form of a finished table (source image w x h)
int build_tiled_plan(sba_ctx* c, const int32_t* lut, int rows, int cols, int w, int h, bool masked, TiledPlan* out);
void free_tiled_plan(TiledPlan* tp);

// ba.cu: the LM solve in three parts (host mailboxes, stream work, collection) so a caller can queue it without waiting
void ba_solve_prepare_host(sba_ba_problem* p, const double* r0, int max_iter);
int ba_solve_enqueue(sba_ba_problem* p, const double t[3], double d1, double d2, double huber, int max_iter, int* launched, bool tran = false);
// the pair pipeline's one-launch variant of ba_solve_enqueue: matched keypoints -> bearings -> the whole LM solve
int ba_pair_solve_enqueue(sba_ba_problem* p, const float* key_l, const float* key_r, const int32_t* qi, const int32_t* ti, const int32_t* d_n,
                          int cap, int cs, int w, int h, const double t[3], double d1, double d2, double huber, int max_iter, int* launched);
int ba_solve_finish(sba_ba_problem* p, double* r_out, const double t[3], double d1, double d2, double huber, int max_iter, int launched,
                    sba_solve_summary* summary, bool tran = false);

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace sba
