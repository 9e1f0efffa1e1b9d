// ba_depth.cu -- depth-only stage of the bundle adjuster and the three-stage solve_problem sequence.
//
// Reference: ba_spherical_costfunctor_d_only (spherical_bundle_adjuster.cpp:1005-1032: three
// reprojection residuals d2*b2 - (R(r)(d1*b1) - t) plus the two barrier residuals lambda*exp(-c*d)),
// its add_residual (:1034-1063: one residual block per match over that match's own 2-vector, no loss
// function, lower bound 0 on both depths, lambda = c = 1) and solve_problem (:183-217: Solve on the
// depth blocks, then the rotation, then the translation).
//
// Every match is its own 2-parameter block, but Ceres solves them as ONE problem: one trust-region
// radius, one accept/reject, one set of convergence tests, and -- because of the bounds -- a projected
// Armijo line search along every valid LM step.  All of that only needs a handful of global sums, so an
// LM trial is ONE pass over the matches:
//
//   per match, in registers (fp64): residual, Jacobian, J^T J (2x2), gradient at x; the damped
//   2x2 Cholesky step for the current radius (exact: every block is eliminated, ITERATIVE_SCHUR's reduced
//   system is empty); the candidate P(x + alpha*step) projected on d >= 0; cost and directional
//   derivative at the candidate.
//   per CTA: ten partial sums/maxima in a fixed order; the last CTA (atomic ticket) folds them in CTA
//   order and runs the whole decision (`depth_decide`: line-search state machine with Ceres' polynomial
//   interpolation, tolerance tests, radius update) -- no host round trip inside a chunk of passes.
//
// HBM traffic per pass and match: b1, b2 (2 x 16 B) + x (16 B) + column scale (16 B) read, candidate
// (16 B) written = 80 B.  x lives in two buffers; accepting a step flips an index instead of copying.
#include "common.cuh"

#include <cfloat>
#include <cmath>

namespace sba {

struct DepthState {
    double cost, radius, dec_factor, alpha, initial_cost;
    double model_dec, gdot, dinf, xnorm2;        // of the LM step being line-searched (from its alpha = 1 pass)
    double prev_x, prev_v, prev_g;               // previous line-search sample
    int prev_valid;
    int cur;                                     // which x buffer holds the current point
    int first;                                   // first pass: project x0, record the Jacobi column scale
    int phase;                                   // 0 = LM step at alpha 1, 1 = line-search trial, 2 = re-evaluate alpha 1 after a failed search
    int ls_iter, ls_evals;
    int iter, num_successful, termination, done, consecutive_invalid, evals, max_iter;
};

struct DepthArgs {
    const float4* b1;
    const float4* b2;
    int n;
    double R[9], t[3], lambda, c;
    double* x[2];
    double* scale;
    double* partial;     // [grid x 10]
    unsigned int* ticket;
    DepthState* st;
};

constexpr int DEPTH_THREADS = 256;
constexpr int DEPTH_SUMS = 10;   // 0 cost_x 1 model_dec 2 gdot 3 xnorm2 4 stepnorm2 5 cost_c 6 dirgrad | max: 7 gradmax 8 dinf 9 bad

// ---- Ceres polynomial.cc / line_search.cc, restated for one device thread ---------------------------

struct LsSample { double x, v, g; };

__device__ inline double poly_eval(const double* p, int deg, double x)
{
    double v = 0;
    for (int k = 0; k <= deg; k++) v = v * x + p[k];
    return v;
}

// FindInterpolatingPolynomial: every sample contributes its value and its gradient; coefficients in
// decreasing powers, degree = 2*ns - 1.
__device__ inline int poly_fit(const LsSample* s, int ns, double* coef)
{
    const int nc = 2 * ns, deg = nc - 1;
    double A[6][7];
    int row = 0;
    for (int i = 0; i < ns; i++) {
        for (int j = 0; j <= deg; j++) A[row][j] = pow(s[i].x, (double)(deg - j));
        A[row][nc] = s[i].v; row++;
        for (int j = 0; j < deg; j++) A[row][j] = (deg - j) * pow(s[i].x, (double)(deg - j - 1));
        A[row][deg] = 0.0;
        A[row][nc] = s[i].g; row++;
    }
    int colperm[6];
    for (int j = 0; j < nc; j++) colperm[j] = j;
    for (int k = 0; k < nc; k++) {   // elimination with full pivoting
        int pr = k, pc = k;
        double best = -1;
        for (int i = k; i < nc; i++)
            for (int j = k; j < nc; j++)
                if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pr = i; pc = j; }
        if (pr != k) for (int j = 0; j <= nc; j++) { double tmp = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = tmp; }
        if (pc != k) {
            for (int i = 0; i < nc; i++) { double tmp = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = tmp; }
            int ti = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = ti;
        }
        if (A[k][k] == 0.0) continue;
        for (int i = k + 1; i < nc; i++) {
            const double f = A[i][k] / A[k][k];
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

// FindPolynomialRoots: real parts of ALL roots (complex pairs included, as Ceres keeps them).
__device__ inline int poly_root_real_parts(const double* p, int deg, double* re)
{
    while (deg > 0 && p[0] == 0.0) { p++; deg--; }
    if (deg == 0) return 0;
    if (deg == 1) { re[0] = -p[1] / p[0]; return 1; }
    if (deg == 2) {
        const double a = p[0], b = p[1], c = p[2], D = b * b - 4 * a * c, sD = sqrt(fabs(D));
        if (D >= 0) {
            if (b >= 0) { re[0] = (-b - sD) / (2.0 * a); re[1] = (2.0 * c) / (-b - sD); }
            else { re[0] = (2.0 * c) / (-b + sD); re[1] = (-b + sD) / (2.0 * a); }
        } else { re[0] = re[1] = -b / (2.0 * a); }
        return 2;
    }
    // simultaneous (Durand-Kerner) iteration on the monic polynomial; Ceres takes companion-matrix eigenvalues
    double m[6], zr[5], zi[5], bound = 0;
    for (int k = 0; k <= deg; k++) m[k] = p[k] / p[0];
    for (int k = 1; k <= deg; k++) bound = fmax(bound, fabs(m[k]));
    bound = 1.0 + bound;
    for (int k = 0; k < deg; k++) {
        const double ang = 2.0 * 3.14159265358979323846 * k / deg + 0.4, rad = bound * (0.5 + 0.5 * (k + 1) / deg);
        zr[k] = rad * cos(ang); zi[k] = rad * sin(ang);
    }
    for (int it = 0; it < 2000; it++) {
        double change = 0;
        for (int k = 0; k < deg; k++) {
            double pr = 1.0, pi = 0.0;
            for (int j = 1; j <= deg; j++) { const double nr = pr * zr[k] - pi * zi[k] + m[j], ni = pr * zi[k] + pi * zr[k]; pr = nr; pi = ni; }
            double qr = 1.0, qi = 0.0;
            for (int j = 0; j < deg; j++)
                if (j != k) {
                    const double dr = zr[k] - zr[j], di = zi[k] - zi[j];
                    const double nr = qr * dr - qi * di, ni = qr * di + qi * dr;
                    qr = nr; qi = ni;
                }
            const double den = qr * qr + qi * qi;
            if (den == 0.0) { zr[k] += 1e-8 * bound; continue; }
            const double wr = (pr * qr + pi * qi) / den, wi = (pi * qr - pr * qi) / den;
            zr[k] -= wr; zi[k] -= wi;
            change = fmax(change, fabs(wr) + fabs(wi));
        }
        if (change <= 1e-15 * bound) break;
    }
    for (int k = 0; k < deg; k++) re[k] = zr[k];
    return deg;
}

// MinimizePolynomial on [x_min, x_max]: mid point, both ends, stationary points.
__device__ inline double poly_minimize(const double* p, int deg, double x_min, double x_max)
{
    double best_x = 0.5 * (x_min + x_max), best_v = poly_eval(p, deg, best_x);
    double v = poly_eval(p, deg, x_min);
    if (v < best_v) { best_v = v; best_x = x_min; }
    v = poly_eval(p, deg, x_max);
    if (v < best_v) { best_v = v; best_x = x_max; }
    if (deg <= 1) return best_x;
    double dp[6], re[6];
    for (int k = 0; k < deg; k++) dp[k] = (deg - k) * p[k];
    const int nr = poly_root_real_parts(dp, deg - 1, re);
    for (int k = 0; k < nr; k++) {
        if (re[k] < x_min || re[k] > x_max) continue;
        v = poly_eval(p, deg, re[k]);
        if (v < best_v) { best_v = v; best_x = re[k]; }
    }
    return best_x;
}

// LineSearch::InterpolatingPolynomialMinimizingStepSize with CUBIC interpolation (the default):
// samples = the start of the search, the current trial and, when there is one, the previous trial.
__device__ inline double ls_next_step(const LsSample& lower, const LsSample& prev, bool prev_valid, const LsSample& cur, double min_step,
                                      double max_step)
{
    if (!isfinite(cur.v)) return fmin(fmax(cur.x * 0.5, min_step), max_step);
    LsSample s[3];
    int ns = 0;
    s[ns++] = lower;
    s[ns++] = cur;
    if (prev_valid) s[ns++] = prev;
    double coef[6];
    const int deg = poly_fit(s, ns, coef);
    return poly_minimize(coef, deg, min_step, max_step);
}

// ---- the decision, one thread -----------------------------------------------------------------------
// Ceres TrustRegionMinimizer (bounds-constrained), LevenbergMarquardtStrategy and ArmijoLineSearch
// defaults; S = the folded sums of the pass that just finished.
__device__ __noinline__ void depth_decide(DepthState& st, const double* S)
{
    const double min_rel_dec = 1e-3, ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, max_radius = 1e16, min_radius = 1e-32;
    const double ls_suff = 1e-4, ls_max_contr = 1e-3, ls_min_contr = 0.6, ls_min_step = 1e-9;
    const int ls_max_iter = 20;
    st.evals++;
    if (st.first) { st.cost = S[0]; st.initial_cost = S[0]; st.first = 0; }

    double cand_cost = S[5], cand_step2 = S[4];
    bool to_post = false;
    if (st.phase == 0) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue, then a new iteration
        if (st.iter >= st.max_iter) { st.termination = 0; st.done = 1; return; }
        if (S[7] <= gtol) { st.termination = 2; st.done = 1; return; }
        if (st.radius < min_radius) { st.termination = 5; st.done = 1; return; }   // Ceres: CONVERGENCE, "minimum trust region radius reached"
        st.iter++;
        if (S[9] != 0.0 || !(S[1] > 0.0)) {   // invalid step: StepIsInvalid
            if (++st.consecutive_invalid >= 5) { st.termination = 4; st.done = 1; return; }
            st.radius *= 0.5;
            return;
        }
        st.consecutive_invalid = 0;
        st.model_dec = S[1]; st.gdot = S[2]; st.xnorm2 = S[3]; st.dinf = S[8];
        st.ls_iter = 0; st.prev_valid = 0;
    } else if (st.phase == 2) {
        to_post = true;   // alpha = 1 re-evaluated after a failed search: take it as it is
    }
    if (!to_post) {
        // ArmijoLineSearch::DoSearch, one trial per pass
        const LsSample cur{st.alpha, S[5], S[6]};
        const bool valid = isfinite(S[5]);
        if (valid && !(S[5] > st.cost + ls_suff * st.gdot * st.alpha)) {
            to_post = true;
        } else {
            bool failed = (++st.ls_iter >= ls_max_iter);
            double next = 0;
            if (!failed) {
                const LsSample lower{0.0, st.cost, st.gdot}, prev{st.prev_x, st.prev_v, st.prev_g};
                next = ls_next_step(lower, prev, st.prev_valid != 0, cur, ls_max_contr * st.alpha, ls_min_contr * st.alpha);
                failed = (next * st.dinf < ls_min_step);
            }
            if (failed) {
                // the step stays as it was (alpha = 1); if the buffers hold another trial, evaluate it again
                if (st.alpha == 1.0) to_post = true;
                else { st.alpha = 1.0; st.phase = 2; return; }
            } else {
                st.prev_x = cur.x; st.prev_v = cur.v; st.prev_g = cur.g; st.prev_valid = valid ? 1 : 0;
                st.alpha = next; st.phase = 1; st.ls_evals++;
                return;
            }
        }
    }
    // ParameterToleranceReached / FunctionToleranceReached on the candidate (not applied on termination)
    if (sqrt(cand_step2) <= ptol * (sqrt(st.xnorm2) + ptol)) { st.termination = 3; st.done = 1; return; }
    const double cost_change = st.cost - cand_cost;
    if (fabs(cost_change) <= ftol * st.cost) { st.termination = 1; st.done = 1; return; }
    const double rel_dec = cost_change / st.model_dec;
    if (rel_dec > min_rel_dec) {
        st.cur ^= 1;
        st.cost = cand_cost;
        st.num_successful++;
        const double q = 2.0 * rel_dec - 1.0;
        st.radius = fmin(max_radius, st.radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
        st.dec_factor = 2.0;
    } else {
        st.radius = st.radius / st.dec_factor;
        st.dec_factor *= 2.0;
    }
    st.alpha = 1.0; st.phase = 0; st.prev_valid = 0;
    if (st.iter >= st.max_iter) { st.termination = 0; st.done = 1; }
}

// ---- one trial pass over the matches ----------------------------------------------------------------

struct DepthPoint {   // everything the pass needs at one point of one match
    double f[3], e0, e1, g0, g1, cost;
};

__device__ inline DepthPoint depth_point(const double u[3], const double b[3], const double t[3], double lambda, double c, double x0, double x1)
{
    DepthPoint P;
    P.e0 = lambda * exp(-c * x0);
    P.e1 = lambda * exp(-c * x1);
    for (int a = 0; a < 3; a++) P.f[a] = x1 * b[a] - (x0 * u[a] - t[a]);
    const double uf = u[0] * P.f[0] + u[1] * P.f[1] + u[2] * P.f[2], bf = b[0] * P.f[0] + b[1] * P.f[1] + b[2] * P.f[2];
    P.g0 = -uf - c * P.e0 * P.e0;     // J^T f with J = [-u, -c e0, 0 ; b, 0, -c e1]
    P.g1 = bf - c * P.e1 * P.e1;
    P.cost = 0.5 * (P.f[0] * P.f[0] + P.f[1] * P.f[1] + P.f[2] * P.f[2] + P.e0 * P.e0 + P.e1 * P.e1);
    return P;
}

__global__ void __launch_bounds__(DEPTH_THREADS, 2) ba_depth_pass_kernel(DepthArgs A)
{
    __shared__ double s_red[DEPTH_THREADS / 32][DEPTH_SUMS];
    __shared__ bool s_last;
    const DepthState& S0 = *A.st;
    if (S0.done) return;
    const int cur = S0.cur, first = S0.first;
    const double radius = S0.radius, alpha = S0.alpha;
    double* __restrict__ xin = A.x[cur];
    double* __restrict__ xout = A.x[cur ^ 1];
    const double min_diag = 1e-6, max_diag = 1e32, inv_radius = 1.0 / radius;

    double acc[DEPTH_SUMS];
#pragma unroll
    for (int k = 0; k < DEPTH_SUMS; k++) acc[k] = 0.0;

    // the loads of the next match are issued before the (long, fp64) arithmetic of this one: two matches' worth of
    // bytes in flight per thread keeps HBM busy at the occupancy this kernel's register count allows
    const int stride = gridDim.x * DEPTH_THREADS;
    int i = blockIdx.x * DEPTH_THREADS + threadIdx.x;
    float4 n1 = make_float4(0.f, 0.f, 0.f, 0.f), n2 = n1;
    double2 nx = make_double2(0.0, 0.0), ns = nx;
    if (i < A.n) {
        n1 = __ldg(A.b1 + i); n2 = __ldg(A.b2 + i);
        nx = reinterpret_cast<const double2*>(xin)[i];
        if (!first) ns = reinterpret_cast<const double2*>(A.scale)[i];
    }
    for (; i < A.n; i += stride) {
        const float4 p1 = n1, p2 = n2;
        double2 xv = nx;
        const double2 sc_loaded = ns;
        if (i + stride < A.n) {
            n1 = __ldg(A.b1 + i + stride); n2 = __ldg(A.b2 + i + stride);
            nx = reinterpret_cast<const double2*>(xin)[i + stride];
            if (!first) ns = reinterpret_cast<const double2*>(A.scale)[i + stride];
        }
        if (first) {   // IterationZero: project the starting point on the box
            xv.x = fmax(xv.x, 0.0); xv.y = fmax(xv.y, 0.0);
            reinterpret_cast<double2*>(xin)[i] = xv;
        }
        const double b[3] = {(double)p2.x, (double)p2.y, (double)p2.z};
        double u[3];
        for (int a = 0; a < 3; a++) u[a] = A.R[3 * a] * (double)p1.x + A.R[3 * a + 1] * (double)p1.y + A.R[3 * a + 2] * (double)p1.z;
        const DepthPoint P = depth_point(u, b, A.t, A.lambda, A.c, xv.x, xv.y);
        const double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2], bb = b[0] * b[0] + b[1] * b[1] + b[2] * b[2],
                     ub = u[0] * b[0] + u[1] * b[1] + u[2] * b[2];
        const double cc = A.c * A.c;
        const double h00 = uu + cc * P.e0 * P.e0, h01 = -ub, h11 = bb + cc * P.e1 * P.e1;
        double2 sc;
        if (first) {
            sc.x = 1.0 / (1.0 + sqrt(h00)); sc.y = 1.0 / (1.0 + sqrt(h11));
            reinterpret_cast<double2*>(A.scale)[i] = sc;
        } else {
            sc = sc_loaded;
        }
        // LevenbergMarquardtStrategy::ComputeStep on the column-scaled block
        const double a00 = h00 * sc.x * sc.x, a01 = h01 * sc.x * sc.y, a11 = h11 * sc.y * sc.y;
        const double gs0 = P.g0 * sc.x, gs1 = P.g1 * sc.y;
        // damped 2x2 system (A + D^2) ds = -gs, D^2 = clamp(diag A) / radius; closed form with ONE reciprocal
        // (fp64 divisions and square roots are what this kernel is bound by); positive definite <=> m00 > 0, det > 0
        const double m00 = a00 + fmin(fmax(a00, min_diag), max_diag) * inv_radius, m11 = a11 + fmin(fmax(a11, min_diag), max_diag) * inv_radius;
        const double det = m00 * m11 - a01 * a01;
        double d0 = 0, d1 = 0;
        if (m00 > 0.0 && det > 0.0) {
            const double inv = 1.0 / det;
            const double ds0 = (a01 * gs1 - m11 * gs0) * inv, ds1 = (a01 * gs0 - m00 * gs1) * inv;
            const double Hd0 = a00 * ds0 + a01 * ds1, Hd1 = a01 * ds0 + a11 * ds1;
            acc[1] -= (gs0 * ds0 + gs1 * ds1) + 0.5 * (ds0 * Hd0 + ds1 * Hd1);
            d0 = ds0 * sc.x; d1 = ds1 * sc.y;
        } else acc[9] = 1.0;
        acc[0] += P.cost;
        acc[2] += P.g0 * d0 + P.g1 * d1;
        acc[3] += xv.x * xv.x + xv.y * xv.y;
        acc[7] = fmax(acc[7], fmax(fabs(xv.x - fmax(xv.x - P.g0, 0.0)), fabs(xv.y - fmax(xv.y - P.g1, 0.0))));
        acc[8] = fmax(acc[8], fmax(fabs(d0), fabs(d1)));
        // candidate: ParameterBlock::Plus projects on the lower bound
        double2 cv;
        cv.x = fmax(xv.x + alpha * d0, 0.0);
        cv.y = fmax(xv.y + alpha * d1, 0.0);
        reinterpret_cast<double2*>(xout)[i] = cv;
        acc[4] += (xv.x - cv.x) * (xv.x - cv.x) + (xv.y - cv.y) * (xv.y - cv.y);
        const DepthPoint Q = depth_point(u, b, A.t, A.lambda, A.c, cv.x, cv.y);
        acc[5] += Q.cost;
        acc[6] += d0 * Q.g0 + d1 * Q.g1;
    }

    // CTA reduction in a fixed order: xor butterflies inside the warp, warps in order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < DEPTH_SUMS; k++) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = (k < 7) ? v + w : fmax(v, w);
        }
        if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < DEPTH_SUMS) {
        const int k = threadIdx.x;
        double v = s_red[0][k];
        for (int w = 1; w < DEPTH_THREADS / 32; w++) v = (k < 7) ? v + s_red[w][k] : fmax(v, s_red[w][k]);
        A.partial[(size_t)blockIdx.x * DEPTH_SUMS + k] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // last CTA: fold the CTA partials in CTA order (thread k owns quantity k; grids are <= a few thousand CTAs)
    __shared__ double s_tot[DEPTH_SUMS];
    {
        // 25 threads per quantity walk the CTAs in strides, then a fixed-order finish by one thread
        __shared__ double s_part[DEPTH_SUMS][25];
        const int k = threadIdx.x / 25, j = threadIdx.x % 25;
        if (k < DEPTH_SUMS) {
            double v = 0.0;
            for (unsigned b = j; b < gridDim.x; b += 25) {
                const double w = __ldcg(A.partial + (size_t)b * DEPTH_SUMS + k);
                v = (k < 7) ? v + w : fmax(v, w);
            }
            s_part[k][j] = v;
        }
        __syncthreads();
        if (threadIdx.x < DEPTH_SUMS) {
            const int q = threadIdx.x;
            double v = s_part[q][0];
            for (int jj = 1; jj < 25; jj++) v = (q < 7) ? v + s_part[q][jj] : fmax(v, s_part[q][jj]);
            s_tot[q] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *A.ticket = 0;
        DepthState st = *A.st;
        depth_decide(st, s_tot);
        *A.st = st;
    }
}

// raw functor values for parity tests: res [n x 5], jac [n x 10] (row-major 5 x 2), cost = 1/2 sum |res|^2 per match
__global__ void ba_depth_functor_kernel(DepthArgs A, const double* __restrict__ d, double* __restrict__ res, double* __restrict__ jac,
                                        double* __restrict__ cost)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n) return;
    const float4 p1 = __ldg(A.b1 + i), p2 = __ldg(A.b2 + i);
    const double b[3] = {(double)p2.x, (double)p2.y, (double)p2.z};
    double u[3];
    for (int a = 0; a < 3; a++) u[a] = A.R[3 * a] * (double)p1.x + A.R[3 * a + 1] * (double)p1.y + A.R[3 * a + 2] * (double)p1.z;
    const DepthPoint P = depth_point(u, b, A.t, A.lambda, A.c, d[2 * i], d[2 * i + 1]);
    if (res) {
        for (int a = 0; a < 3; a++) res[5 * (size_t)i + a] = P.f[a];
        res[5 * (size_t)i + 3] = P.e0;
        res[5 * (size_t)i + 4] = P.e1;
    }
    if (jac) {
        double* J = jac + 10 * (size_t)i;
        for (int a = 0; a < 3; a++) { J[2 * a] = -u[a]; J[2 * a + 1] = b[a]; }
        J[6] = -A.c * P.e0; J[7] = 0.0; J[8] = 0.0; J[9] = -A.c * P.e1;
    }
    if (cost) cost[i] = P.cost;
}

// R(r) as ceres::AngleAxisRotatePoint applies it (host, fp64): both branches.
static void rotation_matrix_host(const double r[3], double R[9])
{
    const double theta2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (theta2 > DBL_EPSILON) {
        const double th = std::sqrt(theta2), c = std::cos(th), s = std::sin(th);
        const double w[3] = {r[0] / th, r[1] / th, r[2] / th};
        const double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) R[3 * a + b] = (a == b ? c : 0.0) + s * K[3 * a + b] + (1.0 - c) * w[a] * w[b];
    } else {
        const double K[9] = {1, -r[2], r[1], r[2], 1, -r[0], -r[1], r[0], 1};
        for (int a = 0; a < 9; a++) R[a] = K[a];
    }
}

static int depth_grid(int n, int sm_count)
{
    int g = (n + DEPTH_THREADS - 1) / DEPTH_THREADS;
    const int cap = sm_count * 2;   // persistent: two resident CTAs per SM (114 registers), each thread walks its matches with the next one prefetched
    if (g > cap) g = cap;
    return g < 1 ? 1 : g;
}

constexpr int DEPTH_CHUNK = 8;   // passes enqueued between two looks at the solver state

}  // namespace sba

using namespace sba;

extern "C" {

int sba_ba_d_eval(sba_ba_problem* p, const double r[3], const double t[3], const double* d, double lambda, double c, double* res, double* jac,
                  double* cost, int mem)
{
    SBA_CHECK_ARG(p && r && t && d);
    const BaView V = ba_problem_view(p);
    if (V.n_cam != 1 || V.n_obs_dev) { sba::set_error("the depth-only block is defined for one camera pair (n_cam == 1)"); return SBA_ERR_UNSUPPORTED; }
    sba_ctx* ctx = V.ctx;
    SBA_CUDA(cudaSetDevice(ctx->device));
    const int n = (int)V.n_obs;
    if (n == 0) return SBA_OK;
    DepthArgs A{};
    A.b1 = V.b1; A.b2 = V.b2; A.n = n; A.lambda = lambda; A.c = c;
    rotation_matrix_host(r, A.R);
    for (int k = 0; k < 3; k++) A.t[k] = t[k];
    const double* d_d;
    double *d_res, *d_jac, *d_cost;
    SBA_TRY(stage_in(ctx, d, (size_t)2 * n, mem, SCR_IN0, &d_d));
    SBA_TRY(stage_out(ctx, res, (size_t)5 * n, mem, SCR_OUT0, &d_res));
    SBA_TRY(stage_out(ctx, jac, (size_t)10 * n, mem, SCR_OUT1, &d_jac));
    SBA_TRY(stage_out(ctx, cost, (size_t)n, mem, SCR_OUT2, &d_cost));
    ba_depth_functor_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(A, d_d, d_res, d_jac, d_cost);
    SBA_LAUNCHED(ctx);
    SBA_TRY(copy_out(ctx, res, d_res, (size_t)5 * n, mem));
    SBA_TRY(copy_out(ctx, jac, d_jac, (size_t)10 * n, mem));
    SBA_TRY(copy_out(ctx, cost, d_cost, (size_t)n, mem));
    return finish(ctx, mem);
}

int sba_ba_d_solve(sba_ba_problem* p, const double r[3], const double t[3], double* d_inout, double lambda, double c, int max_iter,
                   sba_solve_summary* summary, int* line_search_trials, int mem)
{
    SBA_CHECK_ARG(p && r && t && d_inout && max_iter >= 0);
    const BaView V = ba_problem_view(p);
    if (V.n_cam != 1 || V.n_obs_dev) { sba::set_error("the depth-only block is defined for one camera pair (n_cam == 1)"); return SBA_ERR_UNSUPPORTED; }
    sba_ctx* ctx = V.ctx;
    SBA_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int n = (int)V.n_obs;
    const size_t xbytes = (size_t)2 * n * sizeof(double);
    const int grid = depth_grid(n, ctx->sm_count);

    DepthArgs A{};
    A.b1 = V.b1; A.b2 = V.b2; A.n = n; A.lambda = lambda; A.c = c;
    rotation_matrix_host(r, A.R);
    for (int k = 0; k < 3; k++) A.t[k] = t[k];
    DepthState* h_state = nullptr;
    sba::BlockCache& C = ctx->cache;
    int rc = SBA_OK;
    void* blocks[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    auto release = [&]() {
        for (void* b : blocks) C.put(b, false);
        C.put(h_state, true);
    };
#define D_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { sba::set_error(cudaGetErrorString(e_)); release(); return SBA_ERR_CUDA; } } while (0)
    D_CUDA(C.get(&blocks[0], xbytes ? xbytes : 16, false));
    D_CUDA(C.get(&blocks[1], xbytes ? xbytes : 16, false));
    D_CUDA(C.get(&blocks[2], xbytes ? xbytes : 16, false));
    D_CUDA(C.get(&blocks[3], (size_t)grid * DEPTH_SUMS * sizeof(double), false));
    D_CUDA(C.get(&blocks[4], sizeof(unsigned int), false));
    D_CUDA(C.get(&blocks[5], sizeof(DepthState), false));
    D_CUDA(C.get((void**)&h_state, sizeof(DepthState), true));
    A.x[0] = (double*)blocks[0]; A.x[1] = (double*)blocks[1]; A.scale = (double*)blocks[2];
    A.partial = (double*)blocks[3]; A.ticket = (unsigned int*)blocks[4]; A.st = (DepthState*)blocks[5];

    DepthState init{};
    init.radius = 1e4; init.dec_factor = 2.0; init.alpha = 1.0; init.first = 1; init.max_iter = max_iter;
    *h_state = init;
    D_CUDA(cudaMemcpyAsync(A.st, h_state, sizeof(DepthState), cudaMemcpyHostToDevice, st));
    D_CUDA(cudaMemsetAsync(A.ticket, 0, sizeof(unsigned int), st));
    if (n > 0)
        D_CUDA(cudaMemcpyAsync(A.x[0], d_inout, xbytes, mem == SBA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));

    // an LM iteration is one pass, plus one per line-search trial: enqueue chunk by chunk until done
    const long max_passes = (long)(max_iter + 1) * 23 + 2;
    long launched = 0;
    for (;;) {
        prof_begin(ctx, SBA_KERNEL_BA_EVAL);
        for (int k = 0; k < DEPTH_CHUNK; k++) {
            ba_depth_pass_kernel<<<grid, DEPTH_THREADS, 0, st>>>(A);
            SBA_LAUNCHED(ctx);
        }
        prof_end(ctx, SBA_KERNEL_BA_EVAL);
        launched += DEPTH_CHUNK;
        D_CUDA(cudaMemcpyAsync(h_state, A.st, sizeof(DepthState), cudaMemcpyDeviceToHost, st));
        D_CUDA(cudaStreamSynchronize(st));
        if (h_state->done || launched >= max_passes) break;
    }
    if (!h_state->done) { sba::set_error("depth solve did not finish within its pass budget"); rc = SBA_ERR_CUDA; }
    if (rc == SBA_OK && n > 0) {
        D_CUDA(cudaMemcpyAsync(d_inout, A.x[h_state->cur], xbytes, mem == SBA_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
        D_CUDA(cudaStreamSynchronize(st));
    }
    if (summary) {
        summary->iterations = h_state->iter;
        summary->num_successful = h_state->num_successful;
        summary->termination = h_state->termination;
        summary->evaluations = h_state->evals;
        summary->initial_cost = h_state->initial_cost;
        summary->final_cost = h_state->cost;
        summary->final_radius = h_state->radius;
    }
    if (line_search_trials) *line_search_trials = h_state->ls_evals;
    release();
#undef D_CUDA
    return rc;
}

// spherical_bundle_adjuster::solve_problem (spherical_bundle_adjuster.cpp:183-217): depth blocks, then the
// rotation, then the translation -- each stage starts from what the previous one left.  The rotation and
// translation stages take init_d[0][0] and init_d[1][0] (AFTER the depth stage) as their two uniform depths,
// exactly the values the reference's add_residual passes (:941-942, :998-999).
int sba_ba_solve_problem(sba_ba_problem* p, double r_inout[3], double t_inout[3], double* d_inout, double huber_delta, int max_iter,
                         sba_solve_summary summaries[3], int mem)
{
    SBA_CHECK_ARG(p && r_inout && t_inout && d_inout);
    const BaView V = ba_problem_view(p);
    if (V.n_cam != 1 || V.n_obs_dev) { sba::set_error("solve_problem is defined for one camera pair (n_cam == 1)"); return SBA_ERR_UNSUPPORTED; }
    if (V.n_obs < 2) { sba::set_error("solve_problem reads init_d[0][0] and init_d[1][0]: needs at least two matches"); return SBA_ERR_INVALID; }
    sba_solve_summary local[3];
    sba_solve_summary* S = summaries ? summaries : local;
    SBA_TRY(sba_ba_d_solve(p, r_inout, t_inout, d_inout, 1.0, 1.0, max_iter, &S[0], nullptr, mem));
    double d12[3] = {0, 0, 0};   // d[0][0], d[0][1], d[1][0]
    if (mem == SBA_MEM_DEVICE) {
        SBA_CUDA(cudaMemcpyAsync(d12, d_inout, sizeof(d12), cudaMemcpyDeviceToHost, V.ctx->stream));
        SBA_CUDA(cudaStreamSynchronize(V.ctx->stream));
    } else {
        d12[0] = d_inout[0]; d12[2] = d_inout[2];
    }
    SBA_TRY(sba_ba_rot_solve(p, r_inout, t_inout, d12[0], d12[2], huber_delta, max_iter, &S[1]));
    SBA_TRY(sba_ba_tran_solve(p, r_inout, t_inout, d12[0], d12[2], huber_delta, max_iter, &S[2]));
    return SBA_OK;
}

}  // extern "C"
