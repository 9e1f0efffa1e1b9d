/*
 * sba_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity checker for the CUDA product path.  It may be imported /
 * linked / executed only by tests/, __graft_entry__.smoke() and the cpu_baseline /
 * --impl reference legs of bench.py.  Nothing under spherical_bundle_adjuster_b200/
 * links or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * upstream repo whdlgp/spherical_bundle_adjuster).
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - remap / cube2equi : pinned against the REAL reference sources compiled into
 *                         oracle/_ref (oracle/Makefile, cv shim headers) + golden files.
 *   - matcher           : pinned bit-for-bit (indices AND fp32 distances) against
 *                         cv2.BFMatcher(NORM_L2).knnMatch run in the build container;
 *                         golden vectors under tests/golden/.
 *   - BA functor / LM   : PARITY UNPINNED at the Ceres boundary (Ceres is not installed
 *                         and the reference holds no BA test or golden number).  The
 *                         restatement is validated against scipy (Rotation.from_rotvec,
 *                         finite differences, least_squares(loss='huber')) instead.
 *
 * Plain C11, double precision wherever the reference is double.  OpenMP pragmas are only
 * used so the same code can serve as the multi-threaded CPU baseline in bench.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Remap.  equi2cube.cpp:12-280 (six get_* functions) and :282-302 (get_all).
 *
 * Face ids follow the strip order of get_all (equi2cube.cpp:293-298):
 *   0 left, 1 front, 2 right, 3 back, 4 top, 5 bottom.
 * ---------------------------------------------------------------------------------------- */

/* Cartesian direction of face pixel (row i, col j).  equi2cube.cpp:28-30 (back),
 * :73-75 (front), :118-120 (left), :163-165 (right), :208-210 (top), :253-255 (bottom).
 * i and j are doubles so cube2equi_pixel (equi2cube_surf.cpp:22-57) can share it. */
static void face_cart(int face, double i, double j, double cs, double v[3])
{
    switch (face) {
    case 0: /* left   */ v[0] = (cs - 2.0 * j) / cs; v[1] = 1.0;  v[2] = (cs - 2.0 * i) / cs; break;
    case 1: /* front  */ v[0] = -1.0; v[1] = (cs - 2.0 * j) / cs; v[2] = (cs - 2.0 * i) / cs; break;
    case 2: /* right  */ v[0] = (2.0 * j - cs) / cs; v[1] = -1.0; v[2] = (cs - 2.0 * i) / cs; break;
    case 3: /* back   */ v[0] = 1.0;  v[1] = (2.0 * j - cs) / cs; v[2] = (cs - 2.0 * i) / cs; break;
    case 4: /* top    */ v[0] = (cs - 2.0 * i) / cs; v[1] = (cs - 2.0 * j) / cs; v[2] = 1.0;  break;
    default:/* bottom */ v[0] = (2.0 * i - cs) / cs; v[1] = (cs - 2.0 * j) / cs; v[2] = -1.0; break;
    }
}

/* Direction -> (theta, phi).  equi2cube.cpp:32-44. */
static void cart_to_rad(const double v[3], double *theta, double *phi)
{
    double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double ux = v[0] / n, uy = v[1] / n, uz = v[2] / n;
    *theta = acos(uz);
    *phi = atan2(uy, ux);
    if (*phi < 0) *phi += M_PI * 2;
}

/* Source index of one face pixel.  equi2cube.cpp:46-50.  The reference indexes
 * im_data[row*w + col] with no bounds check; row can equal h at the bottom-face centre
 * (theta == pi).  `clamped` (may be NULL) reports whether the clamp fired; the returned index
 * is always in range so the oracle itself never reads out of bounds. */
int32_t orc_equi2cube_src_index(int face, int i, int j, int cs, int w, int h, int *clamped)
{
    double v[3], theta, phi;
    face_cart(face, (double)i, (double)j, (double)cs, v);
    cart_to_rad(v, &theta, &phi);
    int row = (int)(h * theta / M_PI);          /* Vec2i assignment truncates */
    int col = (int)(w * phi / (2 * M_PI));
    int c = 0;
    if (row >= h) { row = h - 1; c = 1; }
    if (col >= w) { col = w - 1; c = 1; }
    if (row < 0) { row = 0; c = 1; }
    if (col < 0) { col = 0; c = 1; }
    if (clamped) *clamped = c;
    return (int32_t)row * w + col;
}

/* Full index table in strip layout: lut[i*(6*cs) + face*cs + j].  Returns #clamped pixels. */
int orc_equi2cube_lut(int cs, int w, int h, int32_t *lut)
{
    int nclamp = 0;
#pragma omp parallel for reduction(+ : nclamp) schedule(static)
    for (int i = 0; i < cs; i++)
        for (int f = 0; f < 6; f++)
            for (int j = 0; j < cs; j++) {
                int c;
                lut[(size_t)i * 6 * cs + (size_t)f * cs + j] = orc_equi2cube_src_index(f, i, j, cs, w, h, &c);
                nclamp += c;
            }
    return nclamp;
}

/* One face, cs x cs x 3 bytes.  equi2cube.cpp:12-55 etc. */
void orc_equi2cube_face(const uint8_t *im, int w, int h, int cs, int face, uint8_t *out)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < cs; i++)
        for (int j = 0; j < cs; j++) {
            int32_t s = orc_equi2cube_src_index(face, i, j, cs, w, h, NULL);
            uint8_t *o = out + ((size_t)i * cs + j) * 3;
            const uint8_t *p = im + (size_t)s * 3;
            o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
        }
}

/* get_all: cs rows x 6cs cols x 3 bytes, faces left,front,right,back,top,bottom.
 * equi2cube.cpp:282-302. */
void orc_equi2cube_all(const uint8_t *im, int w, int h, int cs, uint8_t *strip)
{
#pragma omp parallel for schedule(static)
    for (int i = 0; i < cs; i++)
        for (int f = 0; f < 6; f++)
            for (int j = 0; j < cs; j++) {
                int32_t s = orc_equi2cube_src_index(f, i, j, cs, w, h, NULL);
                uint8_t *o = strip + ((size_t)i * 6 * cs + (size_t)f * cs + j) * 3;
                const uint8_t *p = im + (size_t)s * 3;
                o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
            }
}

/* Strip keypoint -> ERP pixel (float in, double math, float out).
 * equi2cube_surf.cpp:19-76.  xy_in / xy_out are n interleaved (x, y) float pairs. */
void orc_cube2equi_points(const float *xy_in, int n, int cs, int w, int h, float *xy_out)
{
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; k++) {
        float px = xy_in[2 * k], py = xy_in[2 * k + 1];
        double v[3];
        /* Face selection by float comparison against integer multiples of cs
         * (equi2cube_surf.cpp:22,28,34,40,46,52).  x<0 falls in the first branch (left),
         * x>=5cs (including >=6cs) in the last (bottom), exactly as the if/else chain. */
        int face;
        if (px < cs) face = 0;
        else if (px < 2 * cs) face = 1;
        else if (px < 3 * cs) face = 2;
        else if (px < 4 * cs) face = 3;
        else if (px < 5 * cs) face = 4;
        else face = 5;
        /* (cube_pixel.x - k*cube_size) is float - int -> float, then 2.0* promotes. */
        float fx = (face == 0) ? px : (px - (float)(face * cs));
        face_cart(face, (double)py, (double)fx, (double)cs, v);
        double theta, phi;
        cart_to_rad(v, &theta, &phi);
        xy_out[2 * k] = (float)(w * phi / (2 * M_PI));
        xy_out[2 * k + 1] = (float)(h * theta / M_PI);
    }
}

/* ERP pixel -> unit bearing, double.  spherical_bundle_adjuster.cpp:271-298.
 * im_width / im_height are doubles there (:271-272) and pt.x is float: pt.x / im_width
 * promotes to double before the divide. */
void orc_pixels_to_bearings(const float *xy, int n, int w, int h, double *xyz)
{
    double dw = (double)w, dh = (double)h;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < n; k++) {
        double lon = 2 * M_PI * ((double)xy[2 * k] / dw);
        double lat = M_PI * ((double)xy[2 * k + 1] / dh);
        xyz[3 * k] = sin(lat) * cos(lon);
        xyz[3 * k + 1] = sin(lat) * sin(lon);
        xyz[3 * k + 2] = cos(lat);
    }
}

/* ------------------------------------------------------------------------------------------
 * Matcher.  feature_matcher.cpp:42-59 with the matcher north_star fixes (BFMatcher, NORM_L2).
 *
 * Third-party arithmetic: OpenCV BFMatcher::knnMatchImpl -> batchDistance -> batchDistL2_ ->
 * normL2Sqr_ (modules/core/src/norm.cpp; reference pins 3.4.2, the container has 4.13 whose
 * arithmetic was probed and reproduced bit-for-bit):
 *   t = a[k]-b[k] in fp32; sixteen fp32 accumulators (4 vectors x 4 lanes), element k of each
 *   16-chunk goes to accumulator k%16 with a SEPARATE multiply and add (SSE baseline, no FMA);
 *   lane-wise ((d0+d1)+d2)+d3; horizontal (s0+s2)+(s1+s3); a scalar tail for dim%16;
 *   distance = sqrtf(sum).
 * kNN keeps the k smallest with ties resolved to the lower train index.
 * ---------------------------------------------------------------------------------------- */
#if defined(__GNUC__)
#define ORC_NOFMA __attribute__((optimize("fp-contract=off")))
#else
#define ORC_NOFMA
#endif

ORC_NOFMA float orc_l2sqr_opencv(const float *a, const float *b, int n)
{
    float acc[16];
    for (int m = 0; m < 16; m++) acc[m] = 0.f;
    int j = 0;
    float d = 0.f;
    if (n >= 16) {
        for (; j <= n - 16; j += 16)
            for (int m = 0; m < 16; m++) {
                float t = a[j + m] - b[j + m];
                float p = t * t;
                acc[m] = acc[m] + p;
            }
        float s[4];
        for (int l = 0; l < 4; l++) s[l] = ((acc[l] + acc[4 + l]) + acc[8 + l]) + acc[12 + l];
        d = (s[0] + s[2]) + (s[1] + s[3]);
    }
    for (; j < n; j++) {
        float t = a[j] - b[j];
        float p = t * t;
        d = d + p;
    }
    return d;
}

/* kNN k=2 over all train rows.  idx/dist are nq x 2; missing neighbours (nt < 2) are written
 * as idx=-1, dist=+inf.  `dist` holds sqrtf(l2sqr) like DMatch::distance. */
void orc_knn2_l2(const float *q, int nq, const float *t, int nt, int dim, int32_t *idx, float *dist)
{
#pragma omp parallel for schedule(dynamic, 16)
    for (int i = 0; i < nq; i++) {
        float d0 = INFINITY, d1 = INFINITY;
        int32_t i0 = -1, i1 = -1;
        const float *a = q + (size_t)i * dim;
        for (int j = 0; j < nt; j++) {
            float d = sqrtf(orc_l2sqr_opencv(a, t + (size_t)j * dim, dim));
            /* strict '<' keeps the earlier (lower) train index on ties */
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
        dist[2 * i] = d0; dist[2 * i + 1] = d1;
    }
}

/* Lowe ratio filter, feature_matcher.cpp:47-56: keep m[0] iff m0.distance < ratio*m1.distance
 * (fp32 product).  Rows with fewer than two neighbours are skipped (the reference would read
 * knn[i][1] out of bounds there; documented guard).  Returns the number of survivors, written
 * in ascending query order. */
ORC_NOFMA int orc_ratio_filter(const int32_t *idx, const float *dist, int nq, float ratio,
                               int32_t *query_idx, int32_t *train_idx, float *out_dist)
{
    int n = 0;
    for (int i = 0; i < nq; i++) {
        if (idx[2 * i] < 0 || idx[2 * i + 1] < 0) continue;
        float thr = ratio * dist[2 * i + 1];
        if (dist[2 * i] < thr) {
            query_idx[n] = i; train_idx[n] = idx[2 * i]; out_dist[n] = dist[2 * i];
            n++;
        }
    }
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Bundle adjustment, rotation-only.  spherical_bundle_adjuster.cpp:892-919 (functor),
 * :921-945 (per-residual constants), :183-217 + :334-338 (solve, options).
 *
 * Third-party arithmetic (Ceres Solver, version unpinned by the reference's CMakeLists.txt:12,
 * absent here -> restated from its published algorithm; PARITY UNPINNED):
 *   - ceres::AngleAxisRotatePoint (rotation.h): theta2 = r.r; if theta2 > DBL_EPSILON:
 *       theta = sqrt(theta2), w = r/theta, out = p cos + (w x p) sin + w (w.p)(1-cos)
 *     else out = p + r x p.
 *   - AutoDiff Jacobian == analytic derivative of the above (checked by finite differences).
 *   - HuberLoss(a): s=|res|^2, b=a^2; s<=b: rho=s,rho'=1 ; else rho=2a sqrt(s)-b, rho'=a/sqrt(s),
 *     rho''<0 -> Corrector with alpha=0: residual and Jacobian scaled by sqrt(rho').
 *   - cost = 1/2 sum rho(s).
 * ---------------------------------------------------------------------------------------- */

static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

/* ceres::AngleAxisRotatePoint restated. */
void orc_angle_axis_rotate_point(const double r[3], const double p[3], double out[3])
{
    double theta2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (theta2 > DBL_EPSILON) {
        double theta = sqrt(theta2), c = cos(theta), s = sin(theta), ith = 1.0 / theta;
        double w[3] = {r[0] * ith, r[1] * ith, r[2] * ith};
        double wxp[3];
        cross3(w, p, wxp);
        double tmp = (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]) * (1.0 - c);
        for (int k = 0; k < 3; k++) out[k] = p[k] * c + wxp[k] * s + w[k] * tmp;
    } else {
        double rxp[3];
        cross3(r, p, rxp);
        for (int k = 0; k < 3; k++) out[k] = p[k] + rxp[k];
    }
}

/* Rotation matrix R(r) and its three partial derivatives dR/dr_k, both branches consistent
 * with what differentiating AngleAxisRotatePoint (Jets) gives.  R row-major 3x3, dR[k] 3x3. */
void orc_rot_and_derivs(const double r[3], double R[9], double dR[3][9])
{
    double theta2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (theta2 > DBL_EPSILON) {
        double th = sqrt(theta2), c = cos(th), s = sin(th);
        double w[3] = {r[0] / th, r[1] / th, r[2] / th};
        double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        /* R = c I + s K + (1-c) w w^T */
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++)
                R[3 * a + b] = (a == b ? c : 0.0) + s * K[3 * a + b] + (1.0 - c) * w[a] * w[b];
        for (int k = 0; k < 3; k++) {
            /* dtheta/dr_k = w_k ; dw_a/dr_k = (delta_ak - w_a w_k)/theta */
            double dw[3];
            for (int a = 0; a < 3; a++) dw[a] = ((a == k ? 1.0 : 0.0) - w[a] * w[k]) / th;
            double dK[9] = {0, -dw[2], dw[1], dw[2], 0, -dw[0], -dw[1], dw[0], 0};
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++)
                    dR[k][3 * a + b] = (a == b ? -s * w[k] : 0.0) + c * w[k] * K[3 * a + b] + s * dK[3 * a + b] +
                                       s * w[k] * w[a] * w[b] + (1.0 - c) * (dw[a] * w[b] + w[a] * dw[b]);
        }
    } else {
        /* out = p + r x p  ->  R = I + [r]x, dR/dr_k = [e_k]x */
        double K[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        for (int a = 0; a < 9; a++) R[a] = K[a];
        R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
        for (int k = 0; k < 3; k++) {
            double e[3] = {k == 0, k == 1, k == 2};
            double G[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
            for (int a = 0; a < 9; a++) dR[k][a] = G[a];
        }
    }
}

/* One residual + Jacobian, raw (no loss).  spherical_bundle_adjuster.cpp:892-919.
 * res[3]; J[9] row-major: J[3*a+k] = d res_a / d r_k. */
void orc_ba_rot_functor(const double b1[3], const double b2[3], const double r[3], const double t[3],
                        double d1, double d2, double res[3], double J[9])
{
    double X1[3] = {b1[0] * d1, b1[1] * d1, b1[2] * d1};
    double X2[3] = {b2[0] * d2, b2[1] * d2, b2[2] * d2};
    double X1r[3];
    orc_angle_axis_rotate_point(r, X1, X1r);
    for (int a = 0; a < 3; a++) res[a] = X2[a] - (X1r[a] - t[a]);
    if (J) {
        double R[9], dR[3][9];
        orc_rot_and_derivs(r, R, dR);
        for (int k = 0; k < 3; k++)
            for (int a = 0; a < 3; a++)
                J[3 * a + k] = -(dR[k][3 * a] * X1[0] + dR[k][3 * a + 1] * X1[1] + dR[k][3 * a + 2] * X1[2]);
    }
}

/* Huber rho(s) and rho'(s) with scale a (ceres::HuberLoss). */
static void huber(double a, double s, double *rho, double *rho1)
{
    double b = a * a;
    if (a <= 0.0) { *rho = s; *rho1 = 1.0; return; }       /* a<=0: no loss (NULL loss function) */
    if (s > b) { double r = sqrt(s); *rho = 2.0 * a * r - b; *rho1 = fmax(DBL_MIN, a / r); }
    else { *rho = s; *rho1 = 1.0; }
}

/* Evaluate all observations.  b1,b2: n x 3 double.  cam: n camera ids (NULL -> all 0).
 * r: n_cam x 3.  Outputs (any may be NULL): res n x 3 and jac n x 9 are the RAW functor values
 * (what CostFunction::Evaluate returns); H n_cam x 6 (xx,xy,xz,yy,yz,zz), g n_cam x 3 and
 * cost n_cam are the Huber-corrected normal-equation blocks sum J~^T J~, sum J~^T r~, 1/2 sum rho.
 * d1/d2 are the uniform depths of spherical_bundle_adjuster.cpp:941-942. */
void orc_ba_rot_eval(const double *b1, const double *b2, const int32_t *cam, int n, const double *r, int n_cam,
                     const double t[3], double d1, double d2, double huber_a,
                     double *res, double *jac, double *H, double *g, double *cost)
{
    if (H) memset(H, 0, sizeof(double) * 6 * n_cam);
    if (g) memset(g, 0, sizeof(double) * 3 * n_cam);
    if (cost) memset(cost, 0, sizeof(double) * n_cam);
    int want_J = (jac || H || g);
#ifdef _OPENMP
    int nth = omp_get_max_threads();
#else
    int nth = 1;
#endif
    /* per-thread private blocks, combined in thread order (deterministic for a fixed nth) */
    double *priv = (double *)calloc((size_t)nth * n_cam * 10, sizeof(double));
#pragma omp parallel num_threads(nth)
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        double *P = priv + (size_t)tid * n_cam * 10;
#pragma omp for schedule(static)
        for (int i = 0; i < n; i++) {
            int c = cam ? cam[i] : 0;
            double rr[3], J[9];
            orc_ba_rot_functor(b1 + 3 * i, b2 + 3 * i, r + 3 * c, t, d1, d2, rr, want_J ? J : NULL);
            if (res) memcpy(res + 3 * i, rr, 3 * sizeof(double));
            if (jac) memcpy(jac + 9 * i, J, 9 * sizeof(double));
            double s = rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2], rho, rho1;
            huber(huber_a, s, &rho, &rho1);
            double *B = P + (size_t)c * 10;
            B[9] += 0.5 * rho;
            if (want_J) {
                /* J~ = sqrt(rho') J, r~ = sqrt(rho') r  ->  J~^T J~ = rho' J^T J */
                double jx[3] = {J[0], J[3], J[6]}, jy[3] = {J[1], J[4], J[7]}, jz[3] = {J[2], J[5], J[8]};
                double xx = 0, xy = 0, xz = 0, yy = 0, yz = 0, zz = 0, gx = 0, gy = 0, gz = 0;
                for (int a = 0; a < 3; a++) {
                    xx += jx[a] * jx[a]; xy += jx[a] * jy[a]; xz += jx[a] * jz[a];
                    yy += jy[a] * jy[a]; yz += jy[a] * jz[a]; zz += jz[a] * jz[a];
                    gx += jx[a] * rr[a]; gy += jy[a] * rr[a]; gz += jz[a] * rr[a];
                }
                B[0] += rho1 * xx; B[1] += rho1 * xy; B[2] += rho1 * xz;
                B[3] += rho1 * yy; B[4] += rho1 * yz; B[5] += rho1 * zz;
                B[6] += rho1 * gx; B[7] += rho1 * gy; B[8] += rho1 * gz;
            }
        }
    }
    for (int tix = 0; tix < nth; tix++)
        for (int c = 0; c < n_cam; c++) {
            const double *B = priv + ((size_t)tix * n_cam + c) * 10;
            if (H) for (int k = 0; k < 6; k++) H[6 * c + k] += B[k];
            if (g) for (int k = 0; k < 3; k++) g[3 * c + k] += B[6 + k];
            if (cost) cost[c] += B[9];
        }
    free(priv);
}

/* Translation-only functor.  spherical_bundle_adjuster.cpp:948-976 (residual identical to the rot-only
 * one, the free block is t: d res / d t = +I) and :978-1002 (constants per residual: r = init_rot,
 * d1 = init_d[0][0], d2 = init_d[1][0], HuberLoss(1.0), one shared 3-vector block).
 * r: n_cam x 3 fixed rotations, tv: n_cam x 3 translations (the parameters).
 * Outputs as in orc_ba_rot_eval; the Jacobian is the identity and is not materialised. */
void orc_ba_tran_eval(const double *b1, const double *b2, const int32_t *cam, int n, const double *r, const double *tv, int n_cam,
                      double d1, double d2, double huber_a, double *res, double *H, double *g, double *cost)
{
    if (H) memset(H, 0, sizeof(double) * 6 * n_cam);
    if (g) memset(g, 0, sizeof(double) * 3 * n_cam);
    if (cost) memset(cost, 0, sizeof(double) * n_cam);
    for (int i = 0; i < n; i++) {
        int c = cam ? cam[i] : 0;
        double rr[3];
        orc_ba_rot_functor(b1 + 3 * i, b2 + 3 * i, r + 3 * c, tv + 3 * c, d1, d2, rr, NULL);
        if (res) memcpy(res + 3 * i, rr, 3 * sizeof(double));
        double s = rr[0] * rr[0] + rr[1] * rr[1] + rr[2] * rr[2], rho, rho1;
        huber(huber_a, s, &rho, &rho1);
        if (cost) cost[c] += 0.5 * rho;
        if (H) { H[6 * c] += rho1; H[6 * c + 3] += rho1; H[6 * c + 5] += rho1; }   /* rho' I^T I */
        if (g) for (int a = 0; a < 3; a++) g[3 * c + a] += rho1 * rr[a];            /* rho' I^T res */
    }
}

/* Solve the symmetric 3x3 system (H + diag(dd)) x = rhs by Cholesky.  Returns 0 on success. */
static int solve3_spd(const double H[6], const double dd[3], const double rhs[3], double x[3])
{
    double a00 = H[0] + dd[0], a01 = H[1], a02 = H[2], a11 = H[3] + dd[1], a12 = H[4], a22 = H[5] + dd[2];
    if (!(a00 > 0.0)) return 1;
    double l00 = sqrt(a00), l10 = a01 / l00, l20 = a02 / l00;
    double t11 = a11 - l10 * l10;
    if (!(t11 > 0.0)) return 1;
    double l11 = sqrt(t11), l21 = (a12 - l20 * l10) / l11;
    double t22 = a22 - l20 * l20 - l21 * l21;
    if (!(t22 > 0.0)) return 1;
    double l22 = sqrt(t22);
    double y0 = rhs[0] / l00, y1 = (rhs[1] - l10 * y0) / l11, y2 = (rhs[2] - l20 * y0 - l21 * y1) / l22;
    x[2] = y2 / l22;
    x[1] = (y1 - l21 * x[2]) / l11;
    x[0] = (y0 - l10 * x[1] - l20 * x[2]) / l00;
    return 0;
}

typedef struct {
    int iterations;            /* LM iterations executed (successful + unsuccessful) */
    int num_successful;
    int termination;           /* 0 max-iter, 1 function tol, 2 gradient tol, 3 parameter tol, 4 failure (5 invalid steps in a row), 5 minimum trust-region radius (Ceres: CONVERGENCE) */
    double initial_cost;
    double final_cost;
    double final_radius;
} orc_lm_summary;

/* Levenberg-Marquardt over all camera blocks as ONE problem (one trust-region radius, one
 * accept/reject), following Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy defaults
 * named at spherical_bundle_adjuster.cpp:334-338 (max 50 iterations; everything else default):
 *   initial_trust_region_radius 1e4, max 1e16, min 1e-32; min_lm_diagonal 1e-6, max 1e32;
 *   jacobi_scaling on: column scale 1/(1+sqrt(diag(J^T J)));  min_relative_decrease 1e-3;
 *   function_tolerance 1e-6, gradient_tolerance 1e-10, parameter_tolerance 1e-8;
 *   radius update: accepted -> radius /= max(1/3, 1-(2 rho-1)^3), decrease_factor=2;
 *                  rejected -> radius /= decrease_factor, decrease_factor *= 2.
 * The linear solve is exact (dense 3x3 Cholesky per block) where the reference asks for
 * ITERATIVE_SCHUR (:335); with one 3-parameter block per camera the reduced system IS this
 * 3x3 block, so the exact solve is the converged limit of that iteration. */
/* which block is free: 0 = rotation (t fixed), 1 = translation (rotation `fixed` fixed) */
static void lm_eval(int mode, const double *b1, const double *b2, const int32_t *cam, int n, const double *x, int n_cam,
                    const double *fixed, double d1, double d2, double huber_a, double *H, double *g, double *c)
{
    if (mode == 0) orc_ba_rot_eval(b1, b2, cam, n, x, n_cam, fixed, d1, d2, huber_a, NULL, NULL, H, g, c);
    else orc_ba_tran_eval(b1, b2, cam, n, fixed, x, n_cam, d1, d2, huber_a, NULL, H, g, c);
}

static void lm_solve(int mode, const double *b1, const double *b2, const int32_t *cam, int n, double *r, int n_cam,
                     const double *t, double d1, double d2, double huber_a, int max_iter, orc_lm_summary *sum);

void orc_ba_rot_solve(const double *b1, const double *b2, const int32_t *cam, int n, double *r, int n_cam,
                      const double t[3], double d1, double d2, double huber_a, int max_iter,
                      orc_lm_summary *sum)
{
    lm_solve(0, b1, b2, cam, n, r, n_cam, t, d1, d2, huber_a, max_iter, sum);
}

/* spherical_bundle_adjuster.cpp:208-209: the same ceres::Solve on the translation block.
 * r: n_cam x 3 fixed rotations; tv: n_cam x 3 translations, updated in place. */
void orc_ba_tran_solve(const double *b1, const double *b2, const int32_t *cam, int n, const double *r, double *tv, int n_cam,
                       double d1, double d2, double huber_a, int max_iter, orc_lm_summary *sum)
{
    lm_solve(1, b1, b2, cam, n, tv, n_cam, r, d1, d2, huber_a, max_iter, sum);
}

static void lm_solve(int mode, const double *b1, const double *b2, const int32_t *cam, int n, double *r, int n_cam,
                     const double *t, double d1, double d2, double huber_a, int max_iter, orc_lm_summary *sum)
{
    const double min_diag = 1e-6, max_diag = 1e32, min_rel_dec = 1e-3;
    const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, max_radius = 1e16, min_radius = 1e-32;
    double radius = 1e4, dec_factor = 2.0;
    int consecutive_invalid = 0;
    int np = 3 * n_cam;
    double *H = malloc(sizeof(double) * 6 * n_cam), *g = malloc(sizeof(double) * np), *c = malloc(sizeof(double) * n_cam);
    double *Hn = malloc(sizeof(double) * 6 * n_cam), *gn = malloc(sizeof(double) * np), *cn = malloc(sizeof(double) * n_cam);
    double *scale = malloc(sizeof(double) * np), *step = malloc(sizeof(double) * np), *xn = malloc(sizeof(double) * np);

    lm_eval(mode, b1, b2, cam, n, r, n_cam, t, d1, d2, huber_a, H, g, c);
    double cost = 0;
    for (int k = 0; k < n_cam; k++) cost += c[k];
    /* Jacobi scaling is computed once from the initial Jacobian (Ceres does the same). */
    for (int k = 0; k < n_cam; k++) {
        scale[3 * k] = 1.0 / (1.0 + sqrt(H[6 * k]));
        scale[3 * k + 1] = 1.0 / (1.0 + sqrt(H[6 * k + 3]));
        scale[3 * k + 2] = 1.0 / (1.0 + sqrt(H[6 * k + 5]));
    }
    sum->initial_cost = cost; sum->iterations = 0; sum->num_successful = 0; sum->termination = 0;

    double gmax = 0;
    for (int k = 0; k < np; k++) gmax = fmax(gmax, fabs(g[k]));
    if (gmax <= gtol) { sum->termination = 2; goto done; }

    for (int it = 0; it < max_iter; it++) {
        sum->iterations = it + 1;
        /* scaled system: Hs = S H S, gs = S g; D^2 = clamp(diag(Hs))/radius */
        double model_dec = 0, step_norm2 = 0, x_norm2 = 0;
        int bad = 0;
        for (int k = 0; k < n_cam; k++) {
            const double *s = scale + 3 * k;
            double Hs[6] = {H[6 * k] * s[0] * s[0], H[6 * k + 1] * s[0] * s[1], H[6 * k + 2] * s[0] * s[2],
                            H[6 * k + 3] * s[1] * s[1], H[6 * k + 4] * s[1] * s[2], H[6 * k + 5] * s[2] * s[2]};
            double gs[3] = {g[3 * k] * s[0], g[3 * k + 1] * s[1], g[3 * k + 2] * s[2]};
            double dd[3] = {fmin(fmax(Hs[0], min_diag), max_diag) / radius,
                            fmin(fmax(Hs[3], min_diag), max_diag) / radius,
                            fmin(fmax(Hs[5], min_diag), max_diag) / radius};
            double rhs[3] = {-gs[0], -gs[1], -gs[2]}, ds[3];
            if (solve3_spd(Hs, dd, rhs, ds)) { bad = 1; break; }
            /* model decrease = -(gs.ds + 1/2 ds^T Hs ds)  (Ceres: -model_residuals.(f + model_residuals/2)) */
            double Hd[3] = {Hs[0] * ds[0] + Hs[1] * ds[1] + Hs[2] * ds[2], Hs[1] * ds[0] + Hs[3] * ds[1] + Hs[4] * ds[2],
                            Hs[2] * ds[0] + Hs[4] * ds[1] + Hs[5] * ds[2]};
            model_dec -= (gs[0] * ds[0] + gs[1] * ds[1] + gs[2] * ds[2]) + 0.5 * (ds[0] * Hd[0] + ds[1] * Hd[1] + ds[2] * Hd[2]);
            for (int a = 0; a < 3; a++) {
                step[3 * k + a] = ds[a] * s[a];
                xn[3 * k + a] = r[3 * k + a] + step[3 * k + a];
                step_norm2 += step[3 * k + a] * step[3 * k + a];
                x_norm2 += r[3 * k + a] * r[3 * k + a];
            }
        }
        if (bad || !(model_dec > 0.0)) {
            /* invalid step (TrustRegionMinimizer::HandleInvalidStep ->
             * LevenbergMarquardtStrategy::StepIsInvalid): radius *= 0.5, at most 5 in a row */
            if (++consecutive_invalid >= 5) { sum->termination = 4; break; }
            radius *= 0.5;
            if (radius < min_radius) { sum->termination = 5; break; }
            continue;
        }
        consecutive_invalid = 0;

        lm_eval(mode, b1, b2, cam, n, xn, n_cam, t, d1, d2, huber_a, Hn, gn, cn);
        double new_cost = 0;
        for (int k = 0; k < n_cam; k++) new_cost += cn[k];

        /* ParameterToleranceReached / FunctionToleranceReached are tested on the CANDIDATE,
         * before it is accepted; on termination the candidate is NOT applied. */
        if (sqrt(step_norm2) <= ptol * (sqrt(x_norm2) + ptol)) { sum->termination = 3; break; }
        double cost_change = cost - new_cost;
        if (fabs(cost_change) <= ftol * cost) { sum->termination = 1; break; }

        double rel_dec = cost_change / model_dec;
        if (rel_dec > min_rel_dec) {
            memcpy(r, xn, sizeof(double) * np);
            memcpy(H, Hn, sizeof(double) * 6 * n_cam);
            memcpy(g, gn, sizeof(double) * np);
            cost = new_cost;
            sum->num_successful++;
            double q = 2.0 * rel_dec - 1.0;
            radius = radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
            radius = fmin(max_radius, radius);
            dec_factor = 2.0;
            gmax = 0;
            for (int k = 0; k < np; k++) gmax = fmax(gmax, fabs(g[k]));
            if (gmax <= gtol) { sum->termination = 2; break; }
        } else {
            radius = radius / dec_factor; dec_factor *= 2.0;
        }
        /* MinTrustRegionRadiusReached (checked between iterations) */
        if (radius < min_radius) { sum->termination = 5; break; }
    }
done:
    sum->final_cost = cost; sum->final_radius = radius;
    free(H); free(g); free(c); free(Hn); free(gn); free(cn); free(scale); free(step); free(xn);
}

/* ------------------------------------------------------------------------------------------
 * Depth-only block.  spherical_bundle_adjuster.cpp:1005-1032 (functor:
 * 3 reprojection residuals + the two barrier residuals lambda*exp(-c*d)), :1034-1063 (one residual
 * block per match over its own 2-vector init_d[i], NO loss function, lower bound 0 on both depths,
 * lambda = c = 1.0, r = init_rot and t = init_tran constants), solved first at :196-197.
 *
 * Ceres pieces restated (PARITY UNPINNED, see the header):
 *   - TrustRegionMinimizer on a bounds-constrained problem: x0 projected on the box, ParameterBlock::Plus
 *     projects x+delta on the box, gradient tolerance tested on |x - Plus(x, -g)|_inf, and each valid LM
 *     step goes through a projected ARMIJO line search (LineSearch defaults: CUBIC interpolation with
 *     value+directional derivative at every trial, sufficient_function_decrease 1e-4,
 *     max_line_search_step_contraction 1e-3, min_line_search_step_contraction 0.6,
 *     min_line_search_step_size 1e-9, max_num_line_search_step_size_iterations 20).  The model cost
 *     change is NOT recomputed for the shortened step.
 *   - polynomial.cc: FindInterpolatingPolynomial (square Vandermonde-type system),
 *     MinimizePolynomial (mid point, both ends, REAL PARTS of all roots of the derivative inside the
 *     interval), FindPolynomialRoots (closed forms for degree 1 and 2, eigenvalues otherwise -- here a
 *     Durand-Kerner iteration, same roots).
 *   - linear solver: every parameter block is touched by exactly one residual block, so all blocks are
 *     eliminated and ITERATIVE_SCHUR's reduced system is empty: the step is the exact back-substitution,
 *     one damped 2x2 Cholesky per match.
 * ---------------------------------------------------------------------------------------- */

/* res[5]; J[10] row-major 5x2: J[2*a+k] = d res_a / d d_k. */
void orc_ba_d_functor(const double b1[3], const double b2[3], const double r[3], const double t[3], const double d[2],
                      double lambda, double c, double res[5], double J[10])
{
    double X1[3] = {b1[0] * d[0], b1[1] * d[0], b1[2] * d[0]};
    double X1r[3], u[3];
    orc_angle_axis_rotate_point(r, X1, X1r);
    for (int a = 0; a < 3; a++) res[a] = b2[a] * d[1] - (X1r[a] - t[a]);
    double e0 = lambda * exp(-c * d[0]), e1 = lambda * exp(-c * d[1]);
    res[3] = e0;
    res[4] = e1;
    if (J) {
        orc_angle_axis_rotate_point(r, b1, u);      /* d X1r / d d0 = R b1 */
        for (int a = 0; a < 3; a++) { J[2 * a] = -u[a]; J[2 * a + 1] = b2[a]; }
        J[6] = -c * e0; J[7] = 0.0;
        J[8] = 0.0;     J[9] = -c * e1;
    }
}

typedef struct { double x, value, gradient; int value_valid, gradient_valid; } orc_sample;

static double poly_eval(const double *p, int deg, double x)
{
    double v = 0;
    for (int k = 0; k <= deg; k++) v = v * x + p[k];
    return v;
}

/* polynomial.cc FindInterpolatingPolynomial: coefficients in decreasing powers; returns the degree. */
static int poly_fit(const orc_sample *s, int ns, double *coef)
{
    int nc = 0;
    for (int i = 0; i < ns; i++) nc += s[i].value_valid + s[i].gradient_valid;
    int deg = nc - 1, row = 0;
    double A[6][7];
    for (int i = 0; i < ns; i++) {
        if (s[i].value_valid) {
            for (int j = 0; j <= deg; j++) A[row][j] = pow(s[i].x, deg - j);
            A[row][nc] = s[i].value; row++;
        }
        if (s[i].gradient_valid) {
            for (int j = 0; j < deg; j++) A[row][j] = (deg - j) * pow(s[i].x, deg - j - 1);
            A[row][deg] = 0.0;
            A[row][nc] = s[i].gradient; row++;
        }
    }
    /* Gaussian elimination with full pivoting (Ceres: a rank-revealing dense solve) */
    int colperm[6];
    for (int j = 0; j < nc; j++) colperm[j] = j;
    for (int k = 0; k < nc; k++) {
        int pr = k, pc = k; double best = -1;
        for (int i = k; i < nc; i++) for (int j = k; j < nc; j++) if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
        if (pr != k) for (int j = 0; j <= nc; j++) { double tmp = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = tmp; }
        if (pc != k) { for (int i = 0; i < nc; i++) { double tmp = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = tmp; } int ti = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = ti; }
        if (A[k][k] == 0.0) continue;
        for (int i = k + 1; i < nc; i++) {
            double f = A[i][k] / A[k][k];
            for (int j = k; j <= nc; j++) A[i][j] -= f * A[k][j];
        }
    }
    double y[6];
    for (int k = nc - 1; k >= 0; k--) {
        double v = A[k][nc];
        for (int j = k + 1; j < nc; j++) v -= A[k][j] * y[j];
        y[k] = (A[k][k] != 0.0) ? v / A[k][k] : 0.0;
    }
    for (int k = 0; k < nc; k++) coef[colperm[k]] = y[k];
    return deg;
}

/* Real parts of all roots of p (degree deg, decreasing powers).  Returns how many were written. */
static int poly_root_real_parts(const double *p_in, int deg_in, double *re)
{
    const double *p = p_in; int deg = deg_in;
    while (deg > 0 && p[0] == 0.0) { p++; deg--; }          /* RemoveLeadingZeros */
    if (deg == 0) return 0;
    if (deg == 1) { re[0] = -p[1] / p[0]; return 1; }
    if (deg == 2) {                                          /* FindQuadraticPolynomialRoots */
        double a = p[0], b = p[1], c = p[2], D = b * b - 4 * a * c, sD = sqrt(fabs(D));
        if (D >= 0) {
            if (b >= 0) { re[0] = (-b - sD) / (2.0 * a); re[1] = (2.0 * c) / (-b - sD); }
            else { re[0] = (2.0 * c) / (-b + sD); re[1] = (-b + sD) / (2.0 * a); }
        } else { re[0] = re[1] = -b / (2.0 * a); }
        return 2;
    }
    /* Durand-Kerner on the monic polynomial, complex arithmetic by hand (degree <= 4 here) */
    double m[8], zr[8], zi[8], bound = 0;
    for (int k = 0; k <= deg; k++) m[k] = p[k] / p[0];
    for (int k = 1; k <= deg; k++) bound = fmax(bound, fabs(m[k]));
    bound = 1.0 + bound;
    for (int k = 0; k < deg; k++) {                          /* spiral of starting points */
        double ang = 2.0 * M_PI * k / deg + 0.4, rad = bound * (0.5 + 0.5 * (k + 1) / deg);
        zr[k] = rad * cos(ang); zi[k] = rad * sin(ang);
    }
    for (int it = 0; it < 2000; it++) {
        double change = 0;
        for (int k = 0; k < deg; k++) {
            double pr = 1.0, pi = 0.0;                       /* p(z_k), Horner */
            for (int j = 1; j <= deg; j++) { double nr = pr * zr[k] - pi * zi[k] + m[j], ni = pr * zi[k] + pi * zr[k]; pr = nr; pi = ni; }
            double qr = 1.0, qi = 0.0;                       /* prod (z_k - z_j) */
            for (int j = 0; j < deg; j++) if (j != k) {
                double dr = zr[k] - zr[j], di = zi[k] - zi[j];
                double nr = qr * dr - qi * di, ni = qr * di + qi * dr; qr = nr; qi = ni;
            }
            double den = qr * qr + qi * qi;
            if (den == 0.0) { zr[k] += 1e-8 * bound; continue; }
            double wr = (pr * qr + pi * qi) / den, wi = (pi * qr - pr * qi) / den;
            zr[k] -= wr; zi[k] -= wi;
            change = fmax(change, fabs(wr) + fabs(wi));
        }
        if (change <= 1e-15 * bound) break;
    }
    for (int k = 0; k < deg; k++) re[k] = zr[k];
    return deg;
}

/* polynomial.cc MinimizePolynomial */
static double poly_minimize(const double *p, int deg, double x_min, double x_max)
{
    double best_x = 0.5 * (x_min + x_max), best_v = poly_eval(p, deg, best_x);
    double v = poly_eval(p, deg, x_min);
    if (v < best_v) { best_v = v; best_x = x_min; }
    v = poly_eval(p, deg, x_max);
    if (v < best_v) { best_v = v; best_x = x_max; }
    if (deg <= 1) return best_x;
    double dp[6], re[6];
    for (int k = 0; k < deg; k++) dp[k] = (deg - k) * p[k];
    int nr = poly_root_real_parts(dp, deg - 1, re);
    for (int k = 0; k < nr; k++) {
        if (re[k] < x_min || re[k] > x_max) continue;
        v = poly_eval(p, deg, re[k]);
        if (v < best_v) { best_v = v; best_x = re[k]; }
    }
    return best_x;
}

/* line_search.cc LineSearch::InterpolatingPolynomialMinimizingStepSize, CUBIC interpolation */
static double ls_next_step(const orc_sample *lower, const orc_sample *prev, const orc_sample *cur, double min_step, double max_step)
{
    if (!cur->value_valid) return fmin(fmax(cur->x * 0.5, min_step), max_step);
    orc_sample s[3]; int ns = 0;
    s[ns++] = *lower; s[ns++] = *cur;
    if (prev->value_valid) s[ns++] = *prev;
    double coef[6];
    int deg = poly_fit(s, ns, coef);
    return poly_minimize(coef, deg, min_step, max_step);
}

/* exported for the unit test of the polynomial machinery */
double orc_ls_next_step(double f0, double g0, double xp, double fp, double gp, int prev_valid, double xc, double fc, double gc,
                        double min_step, double max_step)
{
    orc_sample lo = {0.0, f0, g0, 1, 1}, pv = {xp, fp, gp, prev_valid, prev_valid}, cu = {xc, fc, gc, 1, 1};
    return ls_next_step(&lo, &pv, &cu, min_step, max_step);
}

/* cost (and gradient, if grad != NULL) at a point that is already feasible */
static double d_cost_grad(const double *b1, const double *b2, int n, const double r[3], const double t[3], const double *x,
                          double lambda, double c, double *grad)
{
    double cost = 0;
#pragma omp parallel for reduction(+ : cost) schedule(static)
    for (int i = 0; i < n; i++) {
        double f[5], J[10];
        orc_ba_d_functor(b1 + 3 * i, b2 + 3 * i, r, t, x + 2 * i, lambda, c, f, grad ? J : NULL);
        double s = 0;
        for (int a = 0; a < 5; a++) s += f[a] * f[a];
        cost += 0.5 * s;
        if (grad) for (int k = 0; k < 2; k++) {
            double gk = 0;
            for (int a = 0; a < 5; a++) gk += J[2 * a + k] * f[a];
            grad[2 * i + k] = gk;
        }
    }
    return cost;
}

/* d: n x 2 depths, updated in place.  ls_evals (optional): line-search trial points beyond alpha = 1. */
void orc_ba_d_solve(const double *b1, const double *b2, int n, const double r[3], const double t[3], double *d,
                    double lambda, double c, int max_iter, orc_lm_summary *sum, int *ls_evals)
{
    const double min_diag = 1e-6, max_diag = 1e32, min_rel_dec = 1e-3;
    const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, max_radius = 1e16, min_radius = 1e-32;
    const double ls_suff = 1e-4, ls_max_contr = 1e-3, ls_min_contr = 0.6, ls_min_step = 1e-9;
    const int ls_max_iter = 20;
    double radius = 1e4, dec_factor = 2.0;
    int consecutive_invalid = 0, n_ls = 0;
    const int np = 2 * n;
    double *g = malloc(sizeof(double) * np), *scale = malloc(sizeof(double) * np), *delta = malloc(sizeof(double) * np);
    double *xc = malloc(sizeof(double) * np), *gc = malloc(sizeof(double) * np), *H = malloc(sizeof(double) * 3 * n);

    for (int k = 0; k < np; k++) d[k] = fmax(d[k], 0.0);        /* IterationZero: project on the box */
    double cost = d_cost_grad(b1, b2, n, r, t, d, lambda, c, g);
    sum->initial_cost = cost; sum->iterations = 0; sum->num_successful = 0; sum->termination = 0;
    int fresh = 1, first = 1;

    for (;;) {
        /* FinalizeIterationAndCheckIfMinimizerCanContinue: iteration limit, gradient, radius */
        if (sum->iterations >= max_iter) { sum->termination = 0; break; }
        if (fresh) {
            /* J^T J blocks at x (the Jacobian only changes when x does) */
            for (int i = 0; i < n; i++) {
                double f[5], J[10];
                orc_ba_d_functor(b1 + 3 * i, b2 + 3 * i, r, t, d + 2 * i, lambda, c, f, J);
                double h00 = 0, h01 = 0, h11 = 0;
                for (int a = 0; a < 5; a++) { h00 += J[2 * a] * J[2 * a]; h01 += J[2 * a] * J[2 * a + 1]; h11 += J[2 * a + 1] * J[2 * a + 1]; }
                H[3 * i] = h00; H[3 * i + 1] = h01; H[3 * i + 2] = h11;
                if (first) { scale[2 * i] = 1.0 / (1.0 + sqrt(h00)); scale[2 * i + 1] = 1.0 / (1.0 + sqrt(h11)); }
            }
            first = 0; fresh = 0;
        }
        double gmax = 0;
        for (int k = 0; k < np; k++) gmax = fmax(gmax, fabs(d[k] - fmax(d[k] - g[k], 0.0)));
        if (gmax <= gtol) { sum->termination = 2; break; }
        if (radius < min_radius) { sum->termination = 5; break; }
        sum->iterations++;

        /* LevenbergMarquardtStrategy::ComputeStep on the column-scaled Jacobian, block by block */
        double model_dec = 0, gdot = 0, dinf = 0, x_norm2 = 0;
        int bad = 0;
        for (int i = 0; i < n; i++) {
            const double s0 = scale[2 * i], s1 = scale[2 * i + 1];
            double a00 = H[3 * i] * s0 * s0, a01 = H[3 * i + 1] * s0 * s1, a11 = H[3 * i + 2] * s1 * s1;
            double g0 = g[2 * i] * s0, g1 = g[2 * i + 1] * s1;
            double d0 = fmin(fmax(a00, min_diag), max_diag) / radius, d1 = fmin(fmax(a11, min_diag), max_diag) / radius;
            double m00 = a00 + d0, m11 = a11 + d1;
            if (!(m00 > 0.0)) { bad = 1; break; }
            double l00 = sqrt(m00), l10 = a01 / l00, t11 = m11 - l10 * l10;
            if (!(t11 > 0.0)) { bad = 1; break; }
            double l11 = sqrt(t11);
            double y0 = -g0 / l00, y1 = (-g1 - l10 * y0) / l11;
            double ds1 = y1 / l11, ds0 = (y0 - l10 * ds1) / l00;
            double Hd0 = a00 * ds0 + a01 * ds1, Hd1 = a01 * ds0 + a11 * ds1;
            model_dec -= (g0 * ds0 + g1 * ds1) + 0.5 * (ds0 * Hd0 + ds1 * Hd1);
            delta[2 * i] = ds0 * s0; delta[2 * i + 1] = ds1 * s1;
            gdot += g[2 * i] * delta[2 * i] + g[2 * i + 1] * delta[2 * i + 1];
            dinf = fmax(dinf, fmax(fabs(delta[2 * i]), fabs(delta[2 * i + 1])));
            x_norm2 += d[2 * i] * d[2 * i] + d[2 * i + 1] * d[2 * i + 1];
        }
        if (bad || !(model_dec > 0.0)) {
            if (++consecutive_invalid >= 5) { sum->termination = 4; break; }
            radius *= 0.5;
            continue;
        }
        consecutive_invalid = 0;

        /* TrustRegionMinimizer::DoLineSearch: projected Armijo search along delta, first trial alpha = 1 */
        double alpha = 1.0;
        {
            orc_sample lower = {0.0, cost, gdot, 1, 1}, prev = {0, 0, 0, 0, 0}, cur;
            int ls_it = 0, ok = 1;
            double a = 1.0;
            for (;;) {
                for (int k = 0; k < np; k++) xc[k] = fmax(d[k] + a * delta[k], 0.0);
                double v = d_cost_grad(b1, b2, n, r, t, xc, lambda, c, gc), dg = 0;
                for (int k = 0; k < np; k++) dg += delta[k] * gc[k];
                cur.x = a; cur.value = v; cur.gradient = dg; cur.value_valid = cur.gradient_valid = isfinite(v) ? 1 : 0;
                if (cur.value_valid && !(v > cost + ls_suff * gdot * a)) break;
                if (++ls_it >= ls_max_iter) { ok = 0; break; }
                double na = ls_next_step(&lower, &prev, &cur, ls_max_contr * a, ls_min_contr * a);
                if (na * dinf < ls_min_step) { ok = 0; break; }
                prev = cur; a = na; n_ls++;
            }
            if (ok) alpha = a;
        }

        /* ComputeCandidatePointAndEvaluateCost */
        double step_norm2 = 0;
        for (int k = 0; k < np; k++) { xc[k] = fmax(d[k] + alpha * delta[k], 0.0); step_norm2 += (d[k] - xc[k]) * (d[k] - xc[k]); }
        double new_cost = d_cost_grad(b1, b2, n, r, t, xc, lambda, c, gc);

        if (sqrt(step_norm2) <= ptol * (sqrt(x_norm2) + ptol)) { sum->termination = 3; break; }
        double cost_change = cost - new_cost;
        if (fabs(cost_change) <= ftol * cost) { sum->termination = 1; break; }
        double rel_dec = cost_change / model_dec;
        if (rel_dec > min_rel_dec) {
            memcpy(d, xc, sizeof(double) * np);
            memcpy(g, gc, sizeof(double) * np);
            cost = new_cost; fresh = 1;
            sum->num_successful++;
            double q = 2.0 * rel_dec - 1.0;
            radius = fmin(max_radius, radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
            dec_factor = 2.0;
        } else {
            radius = radius / dec_factor; dec_factor *= 2.0;
        }
    }
    sum->final_cost = cost; sum->final_radius = radius;
    if (ls_evals) *ls_evals = n_ls;
    free(g); free(scale); free(delta); free(xc); free(gc); free(H);
}

/* ------------------------------------------------------------------------------------------
 * spherical_surf front-end geometry.  spherical_surf.cpp:17-45 (eular2rot),
 * :48-77 (rotate_pixel), :79-109 (crop_rotated_image), :111-123 (rotate_keypoint).
 * Pinned against the reference's own spherical_surf.cpp compiled into oracle/_ref.
 * ---------------------------------------------------------------------------------------- */
#include <limits.h>

/* The Euler angles arrive as floats (cv::Vec3f) and <cmath>'s float overloads of sin/cos are the ones
 * picked, so the factors carry float precision widened to double (spherical_surf.cpp:20-38);
 * R = (R_z * R_y) * R_x in double (:42). */
void orc_eular2rot(const float theta[3], double R[9])
{
    double cx = cosf(theta[0]), sx = sinf(theta[0]), cy = cosf(theta[1]), sy = sinf(theta[1]), cz = cosf(theta[2]), sz = sinf(theta[2]);
    double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
    double Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    double Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
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

/* RAD(x) of spherical_surf.hpp:6, narrowed to float by the Vec3f constructor (:88, :113) */
static float pitch_to_float_rad(float pitch_deg) { return (float)(M_PI * (pitch_deg) / 180.0); }

/* double -> int the way the reference's x86 build converts (cvttsd2si): NaN and out-of-range give INT_MIN */
static int trunc_to_int(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}

/* spherical_surf.cpp:48-77.  (row, col) in, (row, col) out. */
void orc_rotate_pixel(int row, int col, const double R[9], int width, int height, int *row_out, int *col_out)
{
    double lat = M_PI * row / height, lon = 2 * M_PI * col / width;
    double v[3] = {sin(lat) * cos(lon), sin(lat) * sin(lon), cos(lat)};
    double q[3];
    for (int a = 0; a < 3; a++) q[a] = R[3 * a] * v[0] + R[3 * a + 1] * v[1] + R[3 * a + 2] * v[2];
    double th = acos(q[2]), ph = atan2(q[1], q[0]);
    if (ph < 0) ph += M_PI * 2;
    *row_out = trunc_to_int(height * th / M_PI);
    *col_out = trunc_to_int(width * ph / (2 * M_PI));
}

/* Source index (row*w + col) of every pixel of the cropped band, -1 where the bounds check of
 * spherical_surf.cpp:100 fails (the reference leaves those output pixels unwritten). */
void orc_crop_rotated_lut(float pitch_deg, int w, int h, int32_t *lut)
{
    float th[3] = {0.f, pitch_to_float_rad(pitch_deg), 0.f};
    double R[9];
    orc_eular2rot(th, R);
    const int rows = h / 4, off = h * 3 / 8;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < w; j++) {
            int r, c;
            orc_rotate_pixel(i + off, j, R, w, h, &r, &c);
            lut[(size_t)i * w + j] = (r >= 0 && c >= 0 && r < h && c < w) ? r * w + c : -1;
        }
}

/* spherical_surf.cpp:79-109.  out: (h/4) x w x 3; unwritten pixels are set to 0 (the reference leaves
 * them uninitialised). */
void orc_crop_rotated_image(const uint8_t *im, int w, int h, float pitch_deg, uint8_t *out)
{
    const int rows = h / 4;
    int32_t *lut = (int32_t *)malloc(sizeof(int32_t) * (size_t)rows * w);
    orc_crop_rotated_lut(pitch_deg, w, h, lut);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < (int64_t)rows * w; p++) {
        if (lut[p] >= 0) memcpy(out + 3 * p, im + 3 * (size_t)lut[p], 3);
        else memset(out + 3 * p, 0, 3);
    }
    free(lut);
}

/* spherical_surf.cpp:111-123.  xy: n keypoints (x, y) in band coordinates, rotated in place
 * (integer truncation of both coordinates before the rotation, integer results stored as float). */
void orc_rotate_keypoints(float pitch_inv_deg, float *xy, int n, int w, int h)
{
    float th[3] = {0.f, pitch_to_float_rad(pitch_inv_deg), 0.f};
    double R[9];
    orc_eular2rot(th, R);
    for (int k = 0; k < n; k++) {
        int offset_i = (int)(xy[2 * k + 1] + (float)(h * 3 / 8));
        int r, c;
        orc_rotate_pixel(offset_i, (int)xy[2 * k], R, w, h, &r, &c);
        xy[2 * k] = (float)c;
        xy[2 * k + 1] = (float)r;
    }
}

/* ------------------------------------------------------------------------------------------
 * Initial guess.  spherical_bundle_adjuster.cpp:47-115 (eight_point_estimation),
 * :117-181 (initial_guess: 80 random quarter-size subsets, two rotation candidates each, the one
 * closest to the others wins), :24-45 (rot2euler), :14-22 (max_vec).
 *
 * Third-party pieces (OpenCV, restated; pinned against cv2.SVDecomp / cv2.decomposeEssentialMat of the
 * installed OpenCV in tests/test_oracle.py up to the sign freedoms of an SVD):
 *   - cv::SVDecomp: singular values in decreasing order, vt rows to match.  Here: one-sided Jacobi
 *     (Hestenes) on the columns.  The reference only uses the LAST row of vt (the null direction, sign
 *     arbitrary) of the n x 9 system, and U diag(w0, w1, 0) Vt of the 3 x 3 one (unique).
 *   - cv::decomposeEssentialMat: SVD; U, Vt negated if their determinant is negative;
 *     R1 = U W Vt, R2 = U W^T Vt with W = [0 1 0; -1 0 0; 0 0 1]; t = third column of U.
 *     {R1, R2} is invariant under the SVD's sign freedoms, their ORDER and the sign of t are not.
 * ---------------------------------------------------------------------------------------- */

/* One-sided Jacobi SVD of A (m x n, n <= 9, row-major): V (n x n, columns = right singular vectors) and
 * singular values, sorted in decreasing order.  A is overwritten by U*diag(w). */
static void jacobi_svd(double *A, int m, int n, double *w, double *V)
{
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double a = 0, b = 0, c = 0;
                for (int i = 0; i < m; i++) { double x = A[i * n + p], y = A[i * n + q]; a += x * x; b += y * y; c += x * y; }
                if (fabs(c) <= 1e-300 || fabs(c) <= DBL_EPSILON * sqrt(a * b)) continue;
                off = fmax(off, fabs(c) / sqrt(a * b));
                double zeta = (b - a) / (2.0 * c);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int i = 0; i < m; i++) {
                    double x = A[i * n + p], y = A[i * n + q];
                    A[i * n + p] = cs * x - sn * y; A[i * n + q] = sn * x + cs * y;
                }
                for (int i = 0; i < n; i++) {
                    double x = V[i * n + p], y = V[i * n + q];
                    V[i * n + p] = cs * x - sn * y; V[i * n + q] = sn * x + cs * y;
                }
            }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < m; i++) s += A[i * n + j] * A[i * n + j]; w[j] = sqrt(s); }
    for (int j = 0; j < n - 1; j++) {       /* selection sort, decreasing */
        int best = j;
        for (int k = j + 1; k < n; k++) if (w[k] > w[best]) best = k;
        if (best != j) {
            double tw = w[j]; w[j] = w[best]; w[best] = tw;
            for (int i = 0; i < m; i++) { double tmp = A[i * n + j]; A[i * n + j] = A[i * n + best]; A[i * n + best] = tmp; }
            for (int i = 0; i < n; i++) { double tmp = V[i * n + j]; V[i * n + j] = V[i * n + best]; V[i * n + best] = tmp; }
        }
    }
}

static double det3(const double *M)
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

static void mul3(const double *A, const double *B, double *C)
{
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { double s = 0; for (int k = 0; k < 3; k++) s += A[3 * i + k] * B[3 * k + j]; C[3 * i + j] = s; }
}

/* full 3x3 SVD M = U diag(w) Vt from the one-sided Jacobi (U completed by a cross product when w2 ~ 0) */
static void svd3(const double *M, double *U, double *w, double *Vt)
{
    double A[9], V[9];
    memcpy(A, M, sizeof(A));
    jacobi_svd(A, 3, 3, w, V);
    for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) U[3 * i + j] = (w[j] > 1e-12 * w[0]) ? A[3 * i + j] / w[j] : 0.0;
    if (!(w[2] > 1e-12 * w[0])) {     /* rank 2: third left vector = u0 x u1 */
        U[2] = U[3] * U[7] - U[6] * U[4];
        U[5] = U[6] * U[1] - U[0] * U[7];
        U[8] = U[0] * U[4] - U[3] * U[1];
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Vt[3 * i + j] = V[3 * j + i];
}

/* spherical_bundle_adjuster.cpp:24-45 */
static void rot2euler(const double *R, float out[3])
{
    float sy = (float)sqrt(R[0] * R[0] + R[3] * R[3]);
    if (!(sy < 1e-6)) {
        out[0] = (float)atan2(R[7], R[8]); out[1] = (float)atan2(-R[6], sy); out[2] = (float)atan2(R[3], R[0]);
    } else {
        out[0] = (float)atan2(-R[5], R[4]); out[1] = (float)atan2(-R[6], sy); out[2] = 0.f;
    }
}

/* spherical_bundle_adjuster.cpp:14-22 applied to the absolute values (:101-104) */
static double max_vec_abs(const float v[3])
{
    float a = fabsf(v[0]), b = fabsf(v[1]), c = fabsf(v[2]);
    if (a > b && a > c) return a;
    else if (b > c) return b;
    return c;
}

/* From the null direction e (9 numbers, any sign) to the outputs of eight_point_estimation (:71-114). */
void orc_essential_to_candidates(const double e[9], float R1_vec[3], float R2_vec[3], float T_vec[3], int *R1_valid, int *R2_valid)
{
    double U[9], w[3], Vt[9], D[9] = {0}, T[9], Ec[9];
    svd3(e, U, w, Vt);
    D[0] = w[0]; D[4] = w[1]; D[8] = 0.0;                     /* w_f.at<double>(0, 2) = 0 (:74) */
    mul3(U, D, T); mul3(T, Vt, Ec);
    svd3(Ec, U, w, Vt);                                        /* decomposeEssentialMat (:80) */
    if (det3(U) < 0) for (int k = 0; k < 9; k++) U[k] = -U[k];
    if (det3(Vt) < 0) for (int k = 0; k < 9; k++) Vt[k] = -Vt[k];
    const double W[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1}, Wt[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    double R1[9], R2[9];
    mul3(U, W, T); mul3(T, Vt, R1);
    mul3(U, Wt, T); mul3(T, Vt, R2);
    rot2euler(R1, R1_vec);
    rot2euler(R2, R2_vec);
    T_vec[0] = (float)U[2]; T_vec[1] = (float)U[5]; T_vec[2] = (float)U[8];
    *R1_valid = max_vec_abs(R1_vec) < 1.57;
    *R2_valid = max_vec_abs(R2_vec) < 1.57;
}

/* Null direction of the n x 9 epipolar system of one subset (:53-69): rows kron(left_i, right_i). */
void orc_eight_point_null(const double *b1, const double *b2, const int32_t *idx, int n, double e[9], double sv[9])
{
    double *A = (double *)malloc(sizeof(double) * 9 * (size_t)n), w[9], V[81];
    for (int i = 0; i < n; i++) {
        const double *l = b1 + 3 * (size_t)(idx ? idx[i] : i), *r = b2 + 3 * (size_t)(idx ? idx[i] : i);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) A[9 * (size_t)i + 3 * a + b] = l[a] * r[b];
    }
    jacobi_svd(A, n, 9, w, V);
    for (int k = 0; k < 9; k++) e[k] = V[k * 9 + 8];
    if (sv) memcpy(sv, w, sizeof(w));
    free(A);
}

static int cmp_double(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }

/* The vote of initial_guess (:160-180) over the collected candidates.  Returns the winning index. */
int orc_vote_rotation(const float *R_arr, int r)
{
    double *dist = (double *)malloc(sizeof(double) * r), *dn = (double *)malloc(sizeof(double) * r);
    for (int i = 0; i < r; i++) {
        for (int j = 0; j < r; j++) {
            float d0 = R_arr[3 * i] - R_arr[3 * j], d1 = R_arr[3 * i + 1] - R_arr[3 * j + 1], d2 = R_arr[3 * i + 2] - R_arr[3 * j + 2];
            dn[j] = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);      /* float arithmetic on Vec3f elements */
        }
        qsort(dn, r, sizeof(double), cmp_double);
        int lo = (int)(r * 0.2), hi = (int)(r * 0.8);
        double s = 0;
        for (int k = lo; k < hi; k++) s += dn[k];
        dist[i] = s / ((hi - lo) * 1.0);
    }
    int best = 0;
    for (int i = 1; i < r; i++) if (dist[i] < dist[best]) best = i;
    free(dist); free(dn);
    return best;
}

/* initial_guess with the subsets given explicitly: idx [n_samples x sample_n] (the reference draws them with
 * std::random_shuffle, :126-137).  cand_R (optional, capacity 2*n_samples x 3) receives the candidate list. */
int orc_initial_guess(const double *b1, const double *b2, const int32_t *idx, int n_samples, int sample_n, float R_out[3], float T_out[3],
                      float *cand_R, int *n_cand)
{
    float *Ra = (float *)malloc(sizeof(float) * 6 * n_samples), *Ta = (float *)malloc(sizeof(float) * 6 * n_samples);
    int r = 0;
    for (int s = 0; s < n_samples; s++) {
        double e[9];
        float R1[3], R2[3], T[3];
        int v1, v2;
        orc_eight_point_null(b1, b2, idx + (size_t)s * sample_n, sample_n, e, NULL);
        orc_essential_to_candidates(e, R1, R2, T, &v1, &v2);
        if (v1) { memcpy(Ra + 3 * r, R1, 12); memcpy(Ta + 3 * r, T, 12); r++; }
        if (v2) { memcpy(Ra + 3 * r, R2, 12); memcpy(Ta + 3 * r, T, 12); r++; }
    }
    if (n_cand) *n_cand = r;
    if (cand_R) memcpy(cand_R, Ra, sizeof(float) * 3 * r);
    int best = -1;
    if (r > 0) {
        best = orc_vote_rotation(Ra, r);
        memcpy(R_out, Ra + 3 * best, 12);
        memcpy(T_out, Ta + 3 * best, 12);
    }
    free(Ra); free(Ta);
    return best;
}
