// pipeline.cu -- the whole hot path for one ERP pair behind ONE C-ABI call.
//
// Mirrors what a caller of the reference does per pair:
//   equi2cube_surf::do_all            (equi2cube_surf.cpp:78-122): get_all x2, match_two_image,
//                                     cube2equi_pixel per keypoint, gather of the matched keypoints
//   spherical_bundle_adjuster::do_bundle_adjustment (spherical_bundle_adjuster.cpp:268-298, :202-203):
//                                     pixel -> bearing, rotation-only solve
// minus SURF detect/describe (non-free OpenCV, stays on the host: the caller passes keypoints and
// descriptors).  Everything between the input copy and the result copy stays on the device, and no stage
// waits for the host: the match count lives in device memory (the bearing kernel and the BA kernels read
// it there), so the host's only round trip is the LM state after each chunk of evaluations.
//
// No CUDA-graph replay: a captured pair bakes the context's scratch addresses into the graph, which a later,
// larger call may reallocate; it also measured slower than the plain stream path (0.45 vs 0.37 ms per pair in
// round 1), so the path was removed instead of guarded.
#include <cstdlib>

#include "common.cuh"
#include "geometry.cuh"

using namespace sba;

namespace sba {

// Everything that identifies one call's stream work.
struct PairKey {
    const void *erp_l, *erp_r, *strip_l, *strip_r, *desc_l, *desc_r, *key_l, *key_r, *qi, *ti, *dist;
    int w, h, cs, n_left, n_right, dim, mem, max_iter;
    float ratio;
    double t[3], d1, d2, huber;
};

struct PairState {   // what enqueue hands to finish
    int32_t *d_qi, *d_ti;
    float* d_dist;
    bool dev_lists;
    bool have_solve;
};

// All stream work of one pair on c->stream, no synchronisation.  *prob: in = an existing problem to
// reuse or NULL; out = the problem the solve was enqueued on (NULL if none).
static int enqueue_pair(sba_ctx* c, const PairKey& a, const double r0[3], sba_ba_problem** prob, int* launched, PairState* ps)
{
    cudaStream_t st = c->stream;
    const int D = SBA_MEM_DEVICE;
    const int mem = a.mem, w = a.w, h = a.h, cube_size = a.cs, n_left = a.n_left, n_right = a.n_right, dim = a.dim;
    const uint8_t* erp_left = (const uint8_t*)a.erp_l;
    const uint8_t* erp_right = (const uint8_t*)a.erp_r;
    uint8_t* strip_left_out = (uint8_t*)a.strip_l;
    uint8_t* strip_right_out = (uint8_t*)a.strip_r;

    // ---- inputs to the device.  Host mode: descriptors and keypoints (8.6 MB at C2) go first on the
    //      compute stream; the two images (44 MB) are uploaded on a second stream and only the remap,
    //      which runs last, waits for them -- the matcher and the bundle adjustment hide the PCIe time.
    const size_t im_bytes = (size_t)w * h * 3, strip_bytes = (size_t)cube_size * 6 * cube_size * 3;
    const uint8_t *d_im0 = nullptr, *d_im1 = nullptr;
    const float *d_desc0, *d_desc1, *d_key0, *d_key1;
    // When the caller wants the strips back in host memory the reference's order is kept instead: images first,
    // remap, strips on their way to the host (equi2cube_surf.cpp:85-94 hands them to the host-side detector), then the match.
    const bool overlap = erp_left && mem == SBA_MEM_HOST && !strip_left_out && !strip_right_out;
    if (erp_left && !overlap) {
        SBA_TRY(stage_in(c, erp_left, im_bytes, mem, SCR_PIPE_IM0, &d_im0));
        SBA_TRY(stage_in(c, erp_right, im_bytes, mem, SCR_PIPE_IM1, &d_im1));
    }
    SBA_TRY(stage_in(c, (const float*)a.desc_l, (size_t)n_left * dim, mem, SCR_PIPE_DESC0, &d_desc0));
    SBA_TRY(stage_in(c, (const float*)a.desc_r, (size_t)n_right * dim, mem, SCR_PIPE_DESC1, &d_desc1));
    SBA_TRY(stage_in(c, (const float*)a.key_l, (size_t)n_left * 2, mem, SCR_PIPE_KEY0, &d_key0));
    SBA_TRY(stage_in(c, (const float*)a.key_r, (size_t)n_right * 2, mem, SCR_PIPE_KEY1, &d_key1));
    // Device-resident pair: nothing on the device consumes the strips, so the remap runs on the side stream NEXT TO the pair
    // solve (one 16-SM cluster) instead of in front of the matcher; one pair at a time that hides the solve's 50 us.
    static const bool no_fork = getenv("SBA_PAIR_NO_FORK") != nullptr;   // measurement switch
    const bool fork_remap = erp_left && mem == SBA_MEM_DEVICE && !no_fork;
    if ((overlap || fork_remap) && !c->copy_stream) {
        SBA_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        SBA_CUDA(cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming));
        SBA_CUDA(cudaEventCreateWithFlags(&c->copy_ev[1], cudaEventDisableTiming));
        SBA_CUDA(cudaEventCreateWithFlags(&c->main_ev, cudaEventDisableTiming));
    }
    if (overlap) {
        SBA_TRY(c->scratch[SCR_PIPE_IM0].ensure(im_bytes, st));
        SBA_TRY(c->scratch[SCR_PIPE_IM1].ensure(im_bytes, st));
        // the staging buffers may still be read by work already queued on the compute stream
        SBA_CUDA(cudaEventRecord(c->main_ev, st));
        SBA_CUDA(cudaStreamWaitEvent(c->copy_stream, c->main_ev, 0));
        SBA_CUDA(cudaMemcpyAsync(c->scratch[SCR_PIPE_IM0].p, erp_left, im_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        SBA_CUDA(cudaEventRecord(c->copy_ev[0], c->copy_stream));
        SBA_CUDA(cudaMemcpyAsync(c->scratch[SCR_PIPE_IM1].p, erp_right, im_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        SBA_CUDA(cudaEventRecord(c->copy_ev[1], c->copy_stream));
        d_im0 = c->scratch[SCR_PIPE_IM0].as<uint8_t>();
        d_im1 = c->scratch[SCR_PIPE_IM1].as<uint8_t>();
    }

    // ---- equi2cube::get_all on both images (the strips feed the host-side detector of the reference;
    //      they stay on the device unless the caller asks for them).  Enqueued now when the images are
    //      already resident, after everything else when their upload is still in flight.
    auto remap_both = [&]() -> int {
        if (!erp_left) return SBA_OK;
        uint8_t *d_s0, *d_s1;
        const bool dev_out = (mem == SBA_MEM_DEVICE);
        if (dev_out && strip_left_out) d_s0 = strip_left_out;
        else { SBA_TRY(c->scratch[SCR_PIPE_STRIP0].ensure(strip_bytes, st)); d_s0 = c->scratch[SCR_PIPE_STRIP0].as<uint8_t>(); }
        if (dev_out && strip_right_out) d_s1 = strip_right_out;
        else { SBA_TRY(c->scratch[SCR_PIPE_STRIP1].ensure(strip_bytes, st)); d_s1 = c->scratch[SCR_PIPE_STRIP1].as<uint8_t>(); }
        if (overlap) SBA_CUDA(cudaStreamWaitEvent(st, c->copy_ev[1], 0));   // recorded after both uploads on the copy stream
        SBA_TRY(equi2cube_pair(c, d_im0, d_im1, w, h, cube_size, d_s0, d_s1));
        if (!dev_out) {
            SBA_TRY(copy_out(c, strip_left_out, d_s0, strip_bytes, mem));
            SBA_TRY(copy_out(c, strip_right_out, d_s1, strip_bytes, mem));
        }
        return SBA_OK;
    };
    if (!overlap && !fork_remap) SBA_TRY(remap_both());

    // ---- feature_matcher::match_two_image.  The match list goes straight into the caller's device
    //      buffers when there are any; the count stays on the device until the very end.
    const size_t nq = (size_t)(n_left > 0 ? n_left : 1);
    SBA_TRY(c->scratch[SCR_PIPE_MATCH].ensure(nq * 12 + 64, st));
    int32_t* s_qi = c->scratch[SCR_PIPE_MATCH].as<int32_t>();
    int32_t* s_ti = s_qi + nq;
    float* s_dist = (float*)(s_ti + nq);
    int32_t* d_n = (int32_t*)(s_dist + nq);
    ps->dev_lists = (mem == SBA_MEM_DEVICE);
    ps->d_qi = (ps->dev_lists && a.qi) ? (int32_t*)a.qi : s_qi;
    ps->d_ti = (ps->dev_lists && a.ti) ? (int32_t*)a.ti : s_ti;
    ps->d_dist = (ps->dev_lists && a.dist) ? (float*)a.dist : s_dist;
    SBA_TRY(sba_knn2_ratio(c, d_desc0, n_left, d_desc1, n_right, dim, a.ratio, ps->d_qi, ps->d_ti, ps->d_dist, d_n, nullptr, nullptr, D,
                           SBA_MATCH_AUTO));
    // the count also lands in the context's pinned mailbox (written by knn2_finalize_kernel), read after the final synchronise

    ps->have_solve = (n_left > 0 && n_right > 0);
    *launched = 0;
    if (fork_remap) SBA_CUDA(cudaEventRecord(c->main_ev, st));   // the side stream's remap starts once the matcher is through
    if (ps->have_solve) {
        // ---- matched keypoints -> bearings (capacity n_left; the kernel stops at the device-side count)
        SBA_TRY(c->scratch[SCR_PIPE_BEAR].ensure((size_t)2 * n_left * 4 * sizeof(float), st));
        float* d_b = c->scratch[SCR_PIPE_BEAR].as<float>();   // [cap] float4 left bearings, [cap] float4 right bearings
        // ---- matched keypoints -> bearings + the whole rotation-only solve in ONE launch (ba_pair_solve_kernel); the
        //      bearings land in d_b, which the problem borrows, so later stages can still read them
        if (!*prob) SBA_TRY(ba_problem_create_impl(c, d_b, d_b + (size_t)4 * n_left, nullptr, n_left, 1, D, /*borrow=*/true, d_n, prob));
        ba_solve_prepare_host(*prob, r0, a.max_iter);   // pinned mailboxes
        SBA_TRY(ba_pair_solve_enqueue(*prob, d_key0, d_key1, ps->d_qi, ps->d_ti, d_n, n_left, cube_size, w, h, a.t, a.d1, a.d2, a.huber, a.max_iter,
                                      launched));
    }
    if (overlap) SBA_TRY(remap_both());
    if (fork_remap) {
        SBA_CUDA(cudaStreamWaitEvent(c->copy_stream, c->main_ev, 0));
        c->stream = c->copy_stream;            // remap_both enqueues on the context's current stream
        const int rc = remap_both();
        c->stream = st;
        SBA_TRY(rc);
        SBA_CUDA(cudaEventRecord(c->copy_ev[0], c->copy_stream));
        SBA_CUDA(cudaStreamWaitEvent(st, c->copy_ev[0], 0));   // join: whatever follows on the main stream sees the strips
    }
    SBA_CUDA(cudaGetLastError());
    return SBA_OK;
}

// Wait for the enqueued work, finish the solve if its first chunk was not enough, hand results back.
static int finish_pair(sba_ctx* c, const PairKey& a, sba_ba_problem* prob, int launched, const PairState& ps, const double r0[3],
                       sba_pair_result* result)
{
    cudaStream_t st = c->stream;
    double r[3] = {r0[0], r0[1], r0[2]};
    sba_solve_summary sum{};
    if (ps.have_solve) SBA_TRY(ba_solve_finish(prob, r, a.t, a.d1, a.d2, a.huber, a.max_iter, launched, &sum));
    else SBA_CUDA(cudaStreamSynchronize(st));
    const int n = c->pinned_i32[0];
    result->n_matches = n;
    if (n > 0 && ps.have_solve) {
        result->rotation[0] = r[0]; result->rotation[1] = r[1]; result->rotation[2] = r[2];
        result->lm_iterations = sum.iterations;
        result->lm_termination = sum.termination;
        result->initial_cost = sum.initial_cost;
        result->final_cost = sum.final_cost;
    }
    if (!ps.dev_lists) {
        SBA_TRY(copy_out(c, (int32_t*)a.qi, (const int32_t*)ps.d_qi, (size_t)n, a.mem));
        SBA_TRY(copy_out(c, (int32_t*)a.ti, (const int32_t*)ps.d_ti, (size_t)n, a.mem));
        SBA_TRY(copy_out(c, (float*)a.dist, (const float*)ps.d_dist, (size_t)n, a.mem));
    }
    return finish(c, a.mem);
}

static PairKey make_key(const uint8_t* erp_left, const uint8_t* erp_right, int w, int h, int cube_size, uint8_t* strip_left_out,
                        uint8_t* strip_right_out, const float* desc_left, int n_left, const float* desc_right, int n_right, int dim,
                        const float* key_left_xy, const float* key_right_xy, float ratio, const double t[3], double d1, double d2,
                        double huber_delta, int max_iter, int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out, int mem)
{
    PairKey key;
    memset(&key, 0, sizeof(key));
    key.erp_l = erp_left; key.erp_r = erp_right; key.strip_l = strip_left_out; key.strip_r = strip_right_out;
    key.desc_l = desc_left; key.desc_r = desc_right; key.key_l = key_left_xy; key.key_r = key_right_xy;
    key.qi = query_idx_out; key.ti = train_idx_out; key.dist = dist_out;
    key.w = w; key.h = h; key.cs = cube_size; key.n_left = n_left; key.n_right = n_right; key.dim = dim; key.mem = mem; key.max_iter = max_iter;
    key.ratio = ratio; key.t[0] = t[0]; key.t[1] = t[1]; key.t[2] = t[2]; key.d1 = d1; key.d2 = d2; key.huber = huber_delta;
    return key;
}

}  // namespace sba

// One pair whose stream work is queued but whose results have not been collected yet.
struct sba_pair_call {
    sba_ctx* c;
    sba::PairKey key;
    sba::PairState ps;
    sba_ba_problem* prob;
    int launched;
    double r0[3];
};

extern "C" {

int sba_pair_rotation_begin(sba_ctx* c, const uint8_t* erp_left, const uint8_t* erp_right, int w, int h, int cube_size, uint8_t* strip_left_out,
                            uint8_t* strip_right_out, const float* desc_left, int n_left, const float* desc_right, int n_right, int dim,
                            const float* key_left_xy, const float* key_right_xy, float ratio, const double r0[3], const double t[3], double d1,
                            double d2, double huber_delta, int max_iter, int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out,
                            int mem, sba_pair_call** call_out)
{
    SBA_CHECK_ARG(c && call_out && w > 0 && h > 0 && cube_size > 0 && n_left >= 0 && n_right >= 0 && max_iter >= 0);
    SBA_CHECK_ARG(desc_left && desc_right && key_left_xy && key_right_xy && r0 && t);
    SBA_CHECK_ARG((erp_left == nullptr) == (erp_right == nullptr));
    *call_out = nullptr;
    if (c->pair_pending) {
        sba::set_error("a pair is already in flight on this context (its staging buffers are in use): call sba_pair_rotation_end first, "
                       "or use one context per pair in flight");
        return SBA_ERR_INVALID;
    }
    SBA_CUDA(cudaSetDevice(c->device));
    sba_pair_call* call = new sba_pair_call();
    call->c = c;
    call->key = make_key(erp_left, erp_right, w, h, cube_size, strip_left_out, strip_right_out, desc_left, n_left, desc_right, n_right, dim,
                         key_left_xy, key_right_xy, ratio, t, d1, d2, huber_delta, max_iter, query_idx_out, train_idx_out, dist_out, mem);
    call->prob = nullptr;
    call->launched = 0;
    for (int k = 0; k < 3; k++) call->r0[k] = r0[k];
    const int status = enqueue_pair(c, call->key, r0, &call->prob, &call->launched, &call->ps);
    if (status != SBA_OK) {
        cudaStreamSynchronize(c->stream);
        if (call->prob) sba_ba_problem_destroy(call->prob);
        delete call;
        return status;
    }
    c->pair_pending = true;
    *call_out = call;
    return SBA_OK;
}

int sba_pair_rotation_end(sba_pair_call* call, sba_pair_result* result)
{
    SBA_CHECK_ARG(call && result);
    sba_ctx* c = call->c;
    SBA_CUDA(cudaSetDevice(c->device));
    memset(result, 0, sizeof(*result));
    for (int k = 0; k < 3; k++) result->rotation[k] = call->r0[k];
    int status = finish_pair(c, call->key, call->prob, call->launched, call->ps, call->r0, result);
    if (status != SBA_OK || call->key.mem == SBA_MEM_DEVICE) cudaStreamSynchronize(c->stream);   // the problem's buffers go back to the cache
    if (call->prob) sba_ba_problem_destroy(call->prob);
    c->pair_pending = false;
    delete call;
    return status;
}

int sba_pair_rotation(sba_ctx* c, const uint8_t* erp_left, const uint8_t* erp_right, int w, int h, int cube_size, uint8_t* strip_left_out,
                      uint8_t* strip_right_out, const float* desc_left, int n_left, const float* desc_right, int n_right, int dim,
                      const float* key_left_xy, const float* key_right_xy, float ratio, const double r0[3], const double t[3], double d1,
                      double d2, double huber_delta, int max_iter, int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out,
                      sba_pair_result* result, int mem)
{
    SBA_CHECK_ARG(result != nullptr);
    sba_pair_call* call = nullptr;
    SBA_TRY(sba_pair_rotation_begin(c, erp_left, erp_right, w, h, cube_size, strip_left_out, strip_right_out, desc_left, n_left, desc_right, n_right,
                                    dim, key_left_xy, key_right_xy, ratio, r0, t, d1, d2, huber_delta, max_iter, query_idx_out, train_idx_out,
                                    dist_out, mem, &call));
    return sba_pair_rotation_end(call, result);
}

}  // extern "C"
