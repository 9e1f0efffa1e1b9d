// spherical.cu -- geometry of the reference's second front-end, spherical_surf.
//
// Replaces spherical_surf.cpp:17-45 (eular2rot), :48-77 (rotate_pixel), :79-109 (crop_rotated_image),
// :111-123 (rotate_keypoint) and the crop part of do_all (:125-153: three pitched crops + the plain
// equatorial band per image).
//
// Same shape as the cube remap: the source pixel of an output pixel depends only on (w, h, pitch), so
// it is computed once into an int32 table (cached per geometry) and a frame is a table-driven gather
// through the shared warp-cooperative gather kernel (remap.cu).  Differences:
//   * the reference bounds-checks the source pixel and leaves the output pixel unwritten when the check
//     fails (:100-103; it can: a pitch of -90 degrees sends two band pixels onto the poles where
//     acos(|z| > 1) is NaN) -- such entries are -1 in the table and the pixel is written as 0;
//   * the rotation factors are float-precision (cos/sin of a cv::Vec3f pick <cmath>'s float overloads)
//     widened to double, which this file reproduces on the host;
//   * rotate_keypoint (:111-123) truncates both keypoint coordinates to integers and pushes them through
//     the SAME mapping the crop used, so inside the band it is a lookup in the crop's own table.
// Table construction runs on the device in fp64; CUDA's sin/cos/acos/atan2 are not correctly rounded,
// so every pixel whose continuous source coordinate lies within 1e-6 of an integer (or whose rotated z
// is within 1e-9 of the acos domain edge) is re-evaluated on the host with glibc and patched: the
// tables are bit-identical to the reference's arithmetic (tests compare with the reference's own
// spherical_surf.cpp compiled into oracle/_ref).
#include <climits>
#include <cmath>
#include <cstring>

#include "common.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace sba {

struct Rot9 { double m[9]; };

// eular2rot (spherical_surf.cpp:17-45) on the host: float-precision factors, (R_z R_y) R_x in double.
static void eular2rot_host(const float theta[3], double R[9])
{
    const double cx = std::cos(theta[0]), sx = std::sin(theta[0]);   // std::cos(float) -> float, as in the reference
    const double cy = std::cos(theta[1]), sy = std::sin(theta[1]);
    const double cz = std::cos(theta[2]), sz = std::sin(theta[2]);
    const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
    const double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    const double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    double T[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += Rz[3 * i + k] * Ry[3 * k + j];
            T[3 * i + j] = s;
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += T[3 * i + k] * Rx[3 * k + j];
            R[3 * i + j] = s;
        }
}

static Rot9 pitch_rotation(float pitch_deg)
{
    // eular2rot(Vec3f(0, RAD(pitch), 0)): RAD is evaluated in double and narrowed by the Vec3f constructor
    const float th[3] = {0.f, (float)(M_PI * (pitch_deg) / 180.0), 0.f};
    Rot9 R;
    eular2rot_host(th, R.m);
    return R;
}

// double -> int like the reference's x86 build (cvttsd2si): NaN and out-of-range give INT_MIN
__host__ __device__ inline int trunc_like_x86(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}

// rotate_pixel up to (not including) the truncation: continuous (row, col) and the rotated z.
__host__ __device__ inline void rotate_pixel_cont(int row, int col, const double* R, int width, int height, double* row_f, double* col_f,
                                                  double* zrot)
{
    const double lat = M_PI * row / height, lon = 2 * M_PI * col / width;
    const double sl = sin(lat), cl = cos(lat), so = sin(lon), co = cos(lon);
    const double v0 = sl * co, v1 = sl * so, v2 = cl;
    const double q0 = R[0] * v0 + R[1] * v1 + R[2] * v2, q1 = R[3] * v0 + R[4] * v1 + R[5] * v2, q2 = R[6] * v0 + R[7] * v1 + R[8] * v2;
    const double th = acos(q2);
    double ph = atan2(q1, q0);
    if (ph < 0) ph += M_PI * 2;
    *row_f = height * th / M_PI;
    *col_f = width * ph / (2 * M_PI);
    *zrot = q2;
}

__host__ __device__ inline bool needs_host(double rf, double cf, double zrot)
{
    const double tol = 1e-6;
    if (!(fabs(zrot) < 1.0 - 1e-9)) return true;                 // acos domain edge (NaN on one side)
    if (!(rf == rf) || !(cf == cf)) return true;
    return fabs(rf - rint(rf)) < tol || fabs(cf - rint(cf)) < tol;
}

static void rotate_pixel_host(int row, int col, const double* R, int w, int h, int* r, int* c)
{
    double rf, cf, z;
    rotate_pixel_cont(row, col, R, w, h, &rf, &cf, &z);
    *r = trunc_like_x86(rf);
    *c = trunc_like_x86(cf);
}

__host__ __device__ inline int32_t checked_index(int r, int c, int w, int h)
{
    return (r >= 0 && c >= 0 && r < h && c < w) ? r * w + c : -1;    // spherical_surf.cpp:100
}

// ---- crop tables ------------------------------------------------------------------------------------
__global__ void crop_lut_build_kernel(Rot9 R, int w, int h, int32_t* __restrict__ lut, int32_t* __restrict__ flagged,
                                      int* __restrict__ n_flagged, int flag_cap)
{
    const int rows = h / 4, off = h * 3 / 8;
    const int64_t total = (int64_t)rows * w;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / w), j = (int)(p - (int64_t)i * w);
        double rf, cf, z;
        rotate_pixel_cont(i + off, j, R.m, w, h, &rf, &cf, &z);
        lut[p] = checked_index(trunc_like_x86(rf), trunc_like_x86(cf), w, h);
        if (needs_host(rf, cf, z)) {
            const int slot = atomicAdd(n_flagged, 1);
            if (slot < flag_cap) flagged[slot] = (int32_t)p;
        }
    }
}

__global__ void crop_lut_patch_kernel(int32_t* __restrict__ lut, const int32_t* __restrict__ where, const int32_t* __restrict__ what, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) lut[where[k]] = what[k];
}

static int get_crop_plan(sba_ctx* c, int w, int h, float pitch_deg, CropPlan** out)
{
    uint32_t bits;
    std::memcpy(&bits, &pitch_deg, 4);
    const auto key = std::make_tuple(w, h, bits);
    auto it = c->crop_plans.find(key);
    if (it != c->crop_plans.end()) { *out = &it->second; return SBA_OK; }

    const int rows = h / 4, off = h * 3 / 8;
    const int64_t total = (int64_t)rows * w;
    SBA_CHECK_ARG(total > 0 && (int64_t)w * h < ((int64_t)1 << 31));
    CropPlan plan;
    plan.w = w; plan.h = h; plan.pitch_deg = pitch_deg;
    const Rot9 R = pitch_rotation(pitch_deg);
    SBA_CUDA(cudaMalloc(&plan.lut, total * sizeof(int32_t)));
    const int flag_cap = (int)total;
    SBA_TRY(c->scratch[SCR_WORK0].ensure((size_t)flag_cap * sizeof(int32_t), c->stream));
    SBA_TRY(c->scratch[SCR_WORK1].ensure(sizeof(int), c->stream));
    int32_t* d_flag = c->scratch[SCR_WORK0].as<int32_t>();
    int* d_cnt = c->scratch[SCR_WORK1].as<int>();
    SBA_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)c->sm_count * 8);
    crop_lut_build_kernel<<<blocks, 256, 0, c->stream>>>(R, w, h, plan.lut, d_flag, d_cnt, flag_cap);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    int cnt = 0;
    SBA_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    const int nflag = std::min(cnt, flag_cap);
    plan.n_patched = nflag;
    if (nflag > 0) {
        std::vector<int32_t> where(nflag), what(nflag);
        SBA_CUDA(cudaMemcpyAsync(where.data(), d_flag, (size_t)nflag * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        SBA_CUDA(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < nflag; k++) {
            const int i = where[k] / w, j = where[k] - i * w;
            int r, cc;
            rotate_pixel_host(i + off, j, R.m, w, h, &r, &cc);
            what[k] = checked_index(r, cc, w, h);
        }
        SBA_TRY(c->scratch[SCR_WORK2].ensure((size_t)nflag * sizeof(int32_t), c->stream));
        int32_t* d_what = c->scratch[SCR_WORK2].as<int32_t>();
        SBA_CUDA(cudaMemcpyAsync(d_what, what.data(), (size_t)nflag * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        crop_lut_patch_kernel<<<(nflag + 255) / 256, 256, 0, c->stream>>>(plan.lut, d_flag, d_what, nflag);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
        SBA_CUDA(cudaStreamSynchronize(c->stream));
    }
    auto ins = c->crop_plans.emplace(key, plan);
    *out = &ins.first->second;
    return SBA_OK;
}

// The four bands spherical_surf::do_all cuts from each image (:137-143): pitch 45, the plain equatorial
// band (im(roi), rows 3h/8 .. 3h/8 + h/4), pitch -45, pitch -90 -- concatenated into one (4*(h/4)) x w table.
__global__ void band_identity_kernel(int32_t* __restrict__ lut, int w, int h)
{
    const int rows = h / 4, off = h * 3 / 8;
    const int64_t total = (int64_t)rows * w;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) lut[p] = (int32_t)(p + (int64_t)off * w);
}

static const float kBandPitch[4] = {45.f, 0.f, -45.f, -90.f};

static int get_band_plan(sba_ctx* c, int w, int h, int32_t** out)
{
    const auto key = std::make_pair(w, h);
    auto it = c->band_plans.find(key);
    if (it != c->band_plans.end()) { *out = it->second; return SBA_OK; }
    const int64_t per = (int64_t)(h / 4) * w;
    int32_t* lut = nullptr;
    SBA_CUDA(cudaMalloc(&lut, 4 * per * sizeof(int32_t)));
    for (int b = 0; b < 4; b++) {
        if (b == 1) {
            band_identity_kernel<<<(int)std::min<int64_t>(ceil_div64(per, 256), (int64_t)c->sm_count * 8), 256, 0, c->stream>>>(lut + per, w, h);
            SBA_LAUNCHED(c);
        } else {
            CropPlan* plan;
            int rc = get_crop_plan(c, w, h, kBandPitch[b], &plan);
            if (rc != SBA_OK) { cudaFree(lut); return rc; }
            SBA_CUDA(cudaMemcpyAsync(lut + b * per, plan->lut, per * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->stream));
        }
    }
    SBA_CUDA(cudaGetLastError());
    TiledPlan tp;
    int rc = build_tiled_plan(c, lut, 4 * (h / 4), w, w, h, true, &tp);
    if (rc != SBA_OK) { cudaFree(lut); return rc; }
    c->band_plans.emplace(key, lut);
    c->band_tiled.emplace(key, tp);
    *out = lut;
    return SBA_OK;
}

// ---- keypoints / pixels ---------------------------------------------------------------------------
// rc_in: (row, col) integer pairs.  Pixels inside the band of a cached crop table read the table; all others
// compute on the device.  Anything the device cannot settle exactly is queued for the host.
__global__ void rotate_pixels_kernel(const int2* __restrict__ rc_in, int n, Rot9 R, int w, int h, const int32_t* __restrict__ lut, int2* __restrict__ rc_out,
                                     int32_t* __restrict__ flagged, int* __restrict__ n_flagged)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int2 in = rc_in[k];
    const int off = h * 3 / 8, i = in.x - off;
    if (lut && i >= 0 && i < h / 4 && in.y >= 0 && in.y < w) {
        const int32_t s = __ldg(lut + (int64_t)i * w + in.y);
        if (s >= 0) { rc_out[k] = make_int2(s / w, s % w); return; }
    }
    double rf, cf, z;
    rotate_pixel_cont(in.x, in.y, R.m, w, h, &rf, &cf, &z);
    rc_out[k] = make_int2(trunc_like_x86(rf), trunc_like_x86(cf));
    if (needs_host(rf, cf, z)) flagged[atomicAdd(n_flagged, 1)] = k;
}

__global__ void patch_pairs_kernel(int2* __restrict__ out, const int32_t* __restrict__ where, const int2* __restrict__ what, int n)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[where[k]] = what[k];
}

// rotate_keypoint's integer truncation (spherical_surf.cpp:116-118) and its float write-back (:120-121)
__global__ void keypoints_to_pixels_kernel(const float2* __restrict__ xy, int n, int h, int2* __restrict__ rc)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float2 p = xy[k];
    rc[k] = make_int2((int)__fadd_rn(p.y, (float)(h * 3 / 8)), (int)p.x);
}
__global__ void pixels_to_keypoints_kernel(const int2* __restrict__ rc, int n, float2* __restrict__ xy)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    xy[k] = make_float2((float)rc[k].y, (float)rc[k].x);
}

// d_in/d_out: device (row, col) pairs.  Leaves the stream synchronised.
static int rotate_pixels_device(sba_ctx* c, const int2* d_in, int n, const Rot9& R, const int32_t* lut, int w, int h, int2* d_out)
{
    SBA_TRY(c->scratch[SCR_WORK3].ensure((size_t)n * sizeof(int32_t), c->stream));
    SBA_TRY(c->scratch[SCR_WORK4].ensure(sizeof(int), c->stream));
    int32_t* d_flag = c->scratch[SCR_WORK3].as<int32_t>();
    int* d_cnt = c->scratch[SCR_WORK4].as<int>();
    SBA_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), c->stream));
    rotate_pixels_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_in, n, R, w, h, lut, d_out, d_flag, d_cnt);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    int cnt = 0;
    SBA_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    if (cnt > 0) {   // settle the queued ones with the host libm
        std::vector<int32_t> where(cnt);
        std::vector<int2> in(n);
        SBA_CUDA(cudaMemcpyAsync(where.data(), d_flag, (size_t)cnt * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        SBA_CUDA(cudaMemcpyAsync(in.data(), d_in, (size_t)n * sizeof(int2), cudaMemcpyDeviceToHost, c->stream));
        SBA_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<int2> what(cnt);
        for (int k = 0; k < cnt; k++) rotate_pixel_host(in[where[k]].x, in[where[k]].y, R.m, w, h, &what[k].x, &what[k].y);
        SBA_TRY(c->scratch[SCR_WORK5].ensure((size_t)cnt * sizeof(int2), c->stream));
        int2* d_what = c->scratch[SCR_WORK5].as<int2>();
        SBA_CUDA(cudaMemcpyAsync(d_what, what.data(), (size_t)cnt * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
        patch_pairs_kernel<<<(cnt + 255) / 256, 256, 0, c->stream>>>(d_out, d_flag, d_what, cnt);
        SBA_LAUNCHED(c);
        SBA_CUDA(cudaGetLastError());
        SBA_CUDA(cudaStreamSynchronize(c->stream));   // `what` goes out of scope
    }
    return SBA_OK;
}

// the crop table of this geometry if one has been built already (never builds one)
static const int32_t* cached_crop_table(sba_ctx* c, int w, int h, float pitch_deg)
{
    uint32_t bits;
    std::memcpy(&bits, &pitch_deg, 4);
    auto it = c->crop_plans.find(std::make_tuple(w, h, bits));
    return (it != c->crop_plans.end()) ? it->second.lut : nullptr;
}

}  // namespace sba

using namespace sba;

extern "C" {

int sba_eular2rot(const float theta[3], double R_out[9])
{
    SBA_CHECK_ARG(theta && R_out);
    eular2rot_host(theta, R_out);
    return SBA_OK;
}

int sba_crop_rotated_lut(sba_ctx* c, int w, int h, float pitch_deg, int32_t* lut_out, int* n_patched, int mem)
{
    SBA_CHECK_ARG(c && w > 0 && h >= 4);
    SBA_CUDA(cudaSetDevice(c->device));
    CropPlan* plan;
    SBA_TRY(get_crop_plan(c, w, h, pitch_deg, &plan));
    if (n_patched) *n_patched = plan->n_patched;
    if (lut_out)
        SBA_CUDA(cudaMemcpyAsync(lut_out, plan->lut, (size_t)(h / 4) * w * sizeof(int32_t),
                                 mem == SBA_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
    return finish(c, mem);
}

int sba_crop_rotated_image(sba_ctx* c, const uint8_t* erp, int w, int h, int n_images, float pitch_deg, uint8_t* out, int mem)
{
    SBA_CHECK_ARG(c && erp && out && w > 0 && h >= 4 && n_images >= 0);
    if (n_images == 0) return SBA_OK;
    SBA_CUDA(cudaSetDevice(c->device));
    CropPlan* plan;
    SBA_TRY(get_crop_plan(c, w, h, pitch_deg, &plan));
    const size_t in_bytes = (size_t)w * h * 3 * n_images, out_bytes = (size_t)(h / 4) * w * 3 * n_images;
    const uint8_t* d_in;
    uint8_t* d_out;
    SBA_TRY(stage_in(c, erp, in_bytes, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, out, out_bytes, mem, SCR_OUT0, &d_out));
    prof_begin(c, SBA_KERNEL_REMAP);
    SBA_TRY(launch_lut_gather(c, d_in, (int64_t)w * h * 3, plan->lut, h / 4, w, d_out, n_images, true));
    prof_end(c, SBA_KERNEL_REMAP);
    SBA_TRY(copy_out(c, out, d_out, out_bytes, mem));
    return finish(c, mem);
}

int sba_spherical_crops(sba_ctx* c, const uint8_t* erp, int w, int h, int n_images, uint8_t* out, int mem)
{
    SBA_CHECK_ARG(c && erp && out && w > 0 && h >= 4 && n_images >= 0);
    if (n_images == 0) return SBA_OK;
    SBA_CUDA(cudaSetDevice(c->device));
    int32_t* lut;
    SBA_TRY(get_band_plan(c, w, h, &lut));
    const size_t in_bytes = (size_t)w * h * 3 * n_images, out_bytes = (size_t)4 * (h / 4) * w * 3 * n_images;
    const uint8_t* d_in;
    uint8_t* d_out;
    SBA_TRY(stage_in(c, erp, in_bytes, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, out, out_bytes, mem, SCR_OUT0, &d_out));
    prof_begin(c, SBA_KERNEL_REMAP);
    SBA_TRY(launch_lut_gather(c, d_in, (int64_t)w * h * 3, lut, 4 * (h / 4), w, d_out, n_images, true, &c->band_tiled[std::make_pair(w, h)], w));
    prof_end(c, SBA_KERNEL_REMAP);
    SBA_TRY(copy_out(c, out, d_out, out_bytes, mem));
    return finish(c, mem);
}

int sba_rotate_pixels(sba_ctx* c, const int32_t* rc_in, int n, float pitch_deg, int w, int h, int32_t* rc_out, int mem)
{
    SBA_CHECK_ARG(c && n >= 0 && w > 0 && h > 0);
    if (n == 0) return SBA_OK;
    SBA_CHECK_ARG(rc_in && rc_out);
    SBA_CUDA(cudaSetDevice(c->device));
    const int32_t* d_in;
    int32_t* d_out;
    SBA_TRY(stage_in(c, rc_in, (size_t)2 * n, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, rc_out, (size_t)2 * n, mem, SCR_OUT0, &d_out));
    SBA_TRY(rotate_pixels_device(c, (const int2*)d_in, n, pitch_rotation(pitch_deg), cached_crop_table(c, w, h, pitch_deg), w, h, (int2*)d_out));
    SBA_TRY(copy_out(c, rc_out, d_out, (size_t)2 * n, mem));
    return finish(c, mem);
}

int sba_rotate_pixels_mat(sba_ctx* c, const int32_t* rc_in, int n, const double R[9], int w, int h, int32_t* rc_out, int mem)
{
    SBA_CHECK_ARG(c && n >= 0 && w > 0 && h > 0 && R);
    if (n == 0) return SBA_OK;
    SBA_CHECK_ARG(rc_in && rc_out);
    SBA_CUDA(cudaSetDevice(c->device));
    Rot9 M;
    std::memcpy(M.m, R, sizeof(M.m));
    const int32_t* d_in;
    int32_t* d_out;
    SBA_TRY(stage_in(c, rc_in, (size_t)2 * n, mem, SCR_IN0, &d_in));
    SBA_TRY(stage_out(c, rc_out, (size_t)2 * n, mem, SCR_OUT0, &d_out));
    SBA_TRY(rotate_pixels_device(c, (const int2*)d_in, n, M, nullptr, w, h, (int2*)d_out));
    SBA_TRY(copy_out(c, rc_out, d_out, (size_t)2 * n, mem));
    return finish(c, mem);
}

int sba_rotate_keypoints(sba_ctx* c, float* xy_inout, int n, float pitch_inv_deg, int w, int h, int mem)
{
    SBA_CHECK_ARG(c && n >= 0 && w > 0 && h > 0);
    if (n == 0) return SBA_OK;
    SBA_CHECK_ARG(xy_inout);
    SBA_CUDA(cudaSetDevice(c->device));
    const float* d_xy_in;
    SBA_TRY(stage_in(c, (const float*)xy_inout, (size_t)2 * n, mem, SCR_IN0, &d_xy_in));
    float* d_xy = const_cast<float*>(d_xy_in);
    SBA_TRY(c->scratch[SCR_WORK0].ensure((size_t)n * sizeof(int2), c->stream));
    SBA_TRY(c->scratch[SCR_WORK1].ensure((size_t)n * sizeof(int2), c->stream));
    int2* d_rc = c->scratch[SCR_WORK0].as<int2>();
    int2* d_rc_out = c->scratch[SCR_WORK1].as<int2>();
    keypoints_to_pixels_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>((const float2*)d_xy, n, h, d_rc);
    SBA_LAUNCHED(c);
    SBA_TRY(rotate_pixels_device(c, d_rc, n, pitch_rotation(pitch_inv_deg), cached_crop_table(c, w, h, pitch_inv_deg), w, h, d_rc_out));
    pixels_to_keypoints_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d_rc_out, n, (float2*)d_xy);
    SBA_LAUNCHED(c);
    SBA_TRY(copy_out(c, xy_inout, (const float*)d_xy, (size_t)2 * n, mem));
    return finish(c, mem);
}

}  // extern "C"
