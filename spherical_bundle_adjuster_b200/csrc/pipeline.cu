// pipeline.cu -- the whole hot path for one ERP pair behind ONE C-ABI call.
//
// Mirrors what a caller of the reference does per pair:
//   equi2cube_surf::do_all            (equi2cube_surf.cpp:78-122): get_all x2, match_two_image,
//                                     cube2equi_pixel per keypoint, gather of the matched keypoints
//   spherical_bundle_adjuster::do_bundle_adjustment (spherical_bundle_adjuster.cpp:268-298, :202-203):
//                                     pixel -> bearing, rotation-only solve
// minus SURF detect/describe (non-free OpenCV, stays on the host: the caller passes keypoints and
// descriptors).  Everything between the input copy and the result copy stays on the device; the only
// host round trips are the match count (sizes the BA problem) and the LM state once per launch chunk.
#include "common.cuh"
#include "geometry.cuh"

using namespace sba;

namespace sba {

// Matched keypoints -> BA bearings in one pass (equi2cube_surf.cpp:96-113 gather + cube2equi_pixel, then
// spherical_bundle_adjuster.cpp:271-298 pixel -> bearing).  The match count is read from device memory
// so the host never waits for it; threads past the count write nothing.
__global__ void pair_points_kernel(const float2* __restrict__ key_l, const float2* __restrict__ key_r, const int32_t* __restrict__ qi,
                                   const int32_t* __restrict__ ti, const int32_t* __restrict__ d_n, int cap, int cs, int w, int h,
                                   float4* __restrict__ b1, float4* __restrict__ b2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = min(cap, *d_n);
    if (i >= n) return;
    const float2 kl = key_l[qi[i]], kr = key_r[ti[i]];
    float ex, ey;
    double x, y, z;
    cube2equi_point(kl.x, kl.y, cs, w, h, &ex, &ey);
    pixel_to_bearing(ex, ey, (double)w, (double)h, &x, &y, &z);
    b1[i] = make_float4((float)x, (float)y, (float)z, 0.f);
    cube2equi_point(kr.x, kr.y, cs, w, h, &ex, &ey);
    pixel_to_bearing(ex, ey, (double)w, (double)h, &x, &y, &z);
    b2[i] = make_float4((float)x, (float)y, (float)z, 0.f);
}

}  // namespace sba

extern "C" {

int sba_pair_rotation(sba_ctx* c, const uint8_t* erp_left, const uint8_t* erp_right, int w, int h, int cube_size, uint8_t* strip_left_out,
                      uint8_t* strip_right_out, const float* desc_left, int n_left, const float* desc_right, int n_right, int dim,
                      const float* key_left_xy, const float* key_right_xy, float ratio, const double r0[3], const double t[3], double d1,
                      double d2, double huber_delta, int max_iter, int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out,
                      sba_pair_result* result, int mem)
{
    SBA_CHECK_ARG(c && result && w > 0 && h > 0 && cube_size > 0 && n_left >= 0 && n_right >= 0);
    SBA_CHECK_ARG(desc_left && desc_right && key_left_xy && key_right_xy && r0 && t);
    SBA_CHECK_ARG((erp_left == nullptr) == (erp_right == nullptr));
    SBA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    memset(result, 0, sizeof(*result));
    result->rotation[0] = r0[0]; result->rotation[1] = r0[1]; result->rotation[2] = r0[2];
    const int D = SBA_MEM_DEVICE;

    // ---- inputs to the device.  Host mode: descriptors and keypoints (8.6 MB at C2) go first on the
    //      compute stream; the two images (44 MB) are uploaded on a second stream and only the remap,
    //      which runs last, waits for them -- the matcher and the bundle adjustment hide the PCIe time.
    const size_t im_bytes = (size_t)w * h * 3, strip_bytes = (size_t)cube_size * 6 * cube_size * 3;
    const uint8_t *d_im0 = nullptr, *d_im1 = nullptr;
    const float *d_desc0, *d_desc1, *d_key0, *d_key1;
    const bool overlap = erp_left && mem == SBA_MEM_HOST;
    if (overlap) {
        // (image uploads are queued below, after the small inputs)
    } else if (erp_left) {
        SBA_TRY(stage_in(c, erp_left, im_bytes, mem, SCR_PIPE_IM0, &d_im0));
        SBA_TRY(stage_in(c, erp_right, im_bytes, mem, SCR_PIPE_IM1, &d_im1));
    }
    SBA_TRY(stage_in(c, desc_left, (size_t)n_left * dim, mem, SCR_PIPE_DESC0, &d_desc0));
    SBA_TRY(stage_in(c, desc_right, (size_t)n_right * dim, mem, SCR_PIPE_DESC1, &d_desc1));
    SBA_TRY(stage_in(c, key_left_xy, (size_t)n_left * 2, mem, SCR_PIPE_KEY0, &d_key0));
    SBA_TRY(stage_in(c, key_right_xy, (size_t)n_right * 2, mem, SCR_PIPE_KEY1, &d_key1));
    if (overlap) {
        if (!c->copy_stream) {
            SBA_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
            SBA_CUDA(cudaEventCreateWithFlags(&c->copy_ev[0], cudaEventDisableTiming));
            SBA_CUDA(cudaEventCreateWithFlags(&c->copy_ev[1], cudaEventDisableTiming));
            SBA_CUDA(cudaEventCreateWithFlags(&c->main_ev, cudaEventDisableTiming));
        }
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
        if (overlap) SBA_CUDA(cudaStreamWaitEvent(st, c->copy_ev[0], 0));
        SBA_TRY(sba_equi2cube(c, d_im0, w, h, 1, cube_size, d_s0, SBA_MEM_DEVICE));
        if (overlap) SBA_CUDA(cudaStreamWaitEvent(st, c->copy_ev[1], 0));
        SBA_TRY(sba_equi2cube(c, d_im1, w, h, 1, cube_size, d_s1, SBA_MEM_DEVICE));
        if (!dev_out) {
            SBA_TRY(copy_out(c, strip_left_out, d_s0, strip_bytes, mem));
            SBA_TRY(copy_out(c, strip_right_out, d_s1, strip_bytes, mem));
        }
        return SBA_OK;
    };
    if (!overlap) SBA_TRY(remap_both());

    // ---- feature_matcher::match_two_image.  The match list goes straight into the caller's device
    //      buffers when there are any; the count stays on the device until the very end.
    const size_t nq = (size_t)(n_left > 0 ? n_left : 1);
    SBA_TRY(c->scratch[SCR_PIPE_MATCH].ensure(nq * 12 + 64, st));
    int32_t* s_qi = c->scratch[SCR_PIPE_MATCH].as<int32_t>();
    int32_t* s_ti = s_qi + nq;
    float* s_dist = (float*)(s_ti + nq);
    int32_t* d_n = (int32_t*)(s_dist + nq);
    const bool dev_lists = (mem == SBA_MEM_DEVICE);
    int32_t* d_qi = (dev_lists && query_idx_out) ? query_idx_out : s_qi;
    int32_t* d_ti = (dev_lists && train_idx_out) ? train_idx_out : s_ti;
    float* d_dist = (dev_lists && dist_out) ? dist_out : s_dist;
    SBA_TRY(sba_knn2_ratio(c, d_desc0, n_left, d_desc1, n_right, dim, ratio, d_qi, d_ti, d_dist, d_n, nullptr, nullptr, D, SBA_MATCH_AUTO));
    SBA_CUDA(cudaMemcpyAsync(c->pinned_i32, d_n, sizeof(int32_t), cudaMemcpyDeviceToHost, st));   // read after the solve's synchronise

    double r[3] = {r0[0], r0[1], r0[2]};
    if (n_left > 0 && n_right > 0) {
        // ---- matched keypoints -> bearings (capacity n_left; the kernel stops at the device-side count)
        SBA_TRY(c->scratch[SCR_PIPE_BEAR].ensure((size_t)2 * n_left * 4 * sizeof(float), st));
        float* d_b = c->scratch[SCR_PIPE_BEAR].as<float>();        // [cap] float4 left bearings, [cap] float4 right bearings
        pair_points_kernel<<<(n_left + 255) / 256, 256, 0, st>>>((const float2*)d_key0, (const float2*)d_key1, d_qi, d_ti, d_n, n_left, cube_size,
                                                                w, h, (float4*)d_b, (float4*)d_b + n_left);
        SBA_LAUNCHED(c);

        // ---- rotation-only bundle adjustment on the bearings in place
        sba_ba_problem* prob = nullptr;
        SBA_TRY(ba_problem_create_impl(c, d_b, d_b + (size_t)4 * n_left, nullptr, n_left, 1, D, /*borrow=*/true, d_n, &prob));
        sba_solve_summary sum;
        int status = sba_ba_rot_solve(prob, r, t, d1, d2, huber_delta, max_iter, &sum);
        sba_ba_problem_destroy(prob);
        SBA_TRY(status);
        result->lm_iterations = sum.iterations;
        result->lm_termination = sum.termination;
        result->initial_cost = sum.initial_cost;
        result->final_cost = sum.final_cost;
    } else {
        SBA_CUDA(cudaStreamSynchronize(st));
    }
    const int n = c->pinned_i32[0];
    result->n_matches = n;
    if (n > 0) { result->rotation[0] = r[0]; result->rotation[1] = r[1]; result->rotation[2] = r[2]; }
    else { result->lm_iterations = 0; result->lm_termination = 0; result->initial_cost = result->final_cost = 0.0; }
    if (!dev_lists) {
        SBA_TRY(copy_out(c, query_idx_out, (const int32_t*)d_qi, (size_t)n, mem));
        SBA_TRY(copy_out(c, train_idx_out, (const int32_t*)d_ti, (size_t)n, mem));
        SBA_TRY(copy_out(c, dist_out, (const float*)d_dist, (size_t)n, mem));
    }
    if (overlap) SBA_TRY(remap_both());
    return finish(c, mem);
}

}  // extern "C"
