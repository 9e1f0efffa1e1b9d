// geometry.cuh -- cube-face / ERP geometry of the reference, shared by host and device code
// (equi2cube.cpp:26-48 and equi2cube_surf.cpp:19-76: same operation order, no FMA contraction where the
// host code rounds separately).
#pragma once
#include <cmath>
#include <cstdint>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace sba {

__host__ __device__ inline void face_cart(int face, double i, double j, double cs, double v[3])
{
    switch (face) {
    case 0: v[0] = (cs - 2.0 * j) / cs; v[1] = 1.0; v[2] = (cs - 2.0 * i) / cs; break;   // left   :118-120
    case 1: v[0] = -1.0; v[1] = (cs - 2.0 * j) / cs; v[2] = (cs - 2.0 * i) / cs; break;  // front  :73-75
    case 2: v[0] = (2.0 * j - cs) / cs; v[1] = -1.0; v[2] = (cs - 2.0 * i) / cs; break;  // right  :163-165
    case 3: v[0] = 1.0; v[1] = (2.0 * j - cs) / cs; v[2] = (cs - 2.0 * i) / cs; break;   // back   :28-30
    case 4: v[0] = (cs - 2.0 * i) / cs; v[1] = (cs - 2.0 * j) / cs; v[2] = 1.0; break;   // top    :208-210
    default: v[0] = (2.0 * i - cs) / cs; v[1] = (cs - 2.0 * j) / cs; v[2] = -1.0; break; // bottom :253-255
    }
}

__host__ __device__ inline double no_fma_norm(const double v[3])
{
#ifdef __CUDA_ARCH__
    // the host code rounds every product and sum separately; keep the device from contracting to FMA
    return sqrt(__dadd_rn(__dadd_rn(__dmul_rn(v[0], v[0]), __dmul_rn(v[1], v[1])), __dmul_rn(v[2], v[2])));
#else
    return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
#endif
}

// Continuous ERP coordinates (row_f, col_f) of a direction: equi2cube.cpp:32-48 before truncation.
__host__ __device__ inline void dir_to_erp(const double v[3], int w, int h, double* row_f, double* col_f)
{
    double n = no_fma_norm(v);
    double ux = v[0] / n, uy = v[1] / n, uz = v[2] / n;
    double theta = acos(uz);
    double phi = atan2(uy, ux);
    if (phi < 0) phi += M_PI * 2;
#ifdef __CUDA_ARCH__
    *row_f = __dmul_rn((double)h, theta) / M_PI;
    *col_f = __dmul_rn((double)w, phi) / (2 * M_PI);
#else
    *row_f = h * theta / M_PI;
    *col_f = w * phi / (2 * M_PI);
#endif
}

__host__ __device__ inline int32_t clamp_index(double row_f, double col_f, int w, int h, int* clamped)
{
    int row = (int)row_f, col = (int)col_f;  // Vec2i assignment truncates toward zero (:46-48)
    int c = 0;
    if (row >= h) { row = h - 1; c = 1; }
    if (col >= w) { col = w - 1; c = 1; }
    if (row < 0) { row = 0; c = 1; }
    if (col < 0) { col = 0; c = 1; }
    if (clamped) *clamped = c;
    return row * w + col;
}


// equi2cube_surf::cube2equi_pixel (equi2cube_surf.cpp:19-76): strip keypoint -> ERP pixel (float out).
__host__ __device__ inline void cube2equi_point(float px, float py, int cs, int w, int h, float* ox, float* oy)
{
    int face;
    if (px < cs) face = 0;
    else if (px < 2 * cs) face = 1;
    else if (px < 3 * cs) face = 2;
    else if (px < 4 * cs) face = 3;
    else if (px < 5 * cs) face = 4;
    else face = 5;
#ifdef __CUDA_ARCH__
    const float fx = (face == 0) ? px : __fsub_rn(px, (float)(face * cs));   // float - int in the reference
#else
    const float fx = (face == 0) ? px : (px - (float)(face * cs));
#endif
    double v[3], rf, cf;
    face_cart(face, (double)py, (double)fx, (double)cs, v);
    dir_to_erp(v, w, h, &rf, &cf);
    *ox = (float)cf;
    *oy = (float)rf;
}

// ERP pixel -> unit bearing (spherical_bundle_adjuster.cpp:271-298), fp64.
__device__ inline void pixel_to_bearing(float x, float y, double w, double h, double* bx, double* by, double* bz)
{
    const double lon = 2 * M_PI * ((double)x / w);
    const double lat = M_PI * ((double)y / h);
    double sl, cl, so, co;
    sincos(lat, &sl, &cl);
    sincos(lon, &so, &co);
    *bx = sl * co; *by = sl * so; *bz = cl;
}

}  // namespace sba
