// ba.cu -- rotation-only spherical bundle adjustment on the device (K5 + K6).
//
// Replaces ba_spherical_costfunctor_rot_only::operator() / add_residual
// (spherical_bundle_adjuster.cpp:892-945) plus the ceres::Solve call that consumes them
// (:183-217, options :334-338).
//
// Data layout in HBM
//   b1, b2   : n_obs x float4 (x, y, z, -), observations grouped by camera at problem creation so a
//              work item (<= 8192 consecutive observations of ONE camera) is a contiguous, 16-byte
//              aligned range; 32 B read per observation, nothing else on the fused path.
//   params   : per camera {d1*R (9 fp64), dR/dr_k (27 fp64), -dR/dr_k (27 fp32)}, rebuilt whenever r changes.
//   partial  : n_items x 16 fp64 {S (6), C (9), cost}: the weighted moments of one work item (below).
//   blocks   : n_cam x 10 fp64 {Hxx,Hxy,Hxz,Hyy,Hyz,Hzz,gx,gy,gz,cost}, the per-camera normal-equation
//              blocks (what an all-reduce sums).
//
// Moment form.  J_i = -dR(r)/dr . X1_i is LINEAR in the observation with per-camera constants, so
//   H = sum_i w_i J_i^T J_i  and  g = sum_i w_i J_i^T res_i
// depend on the observations only through  S = sum_i w_i b1_i b1_i^T (6 unique)  and
// C = sum_i w_i b1_i res_i^T (9)  (w_i = Huber rho').  The evaluation kernel therefore accumulates 16
// scalars per observation in fp64 (B200: 64 DFMA/clk/SM -- the kernel stays HBM-bound) and the
// contraction with dR/dr (fp64, ~200 flops) happens once per camera when the partials are folded.
// Everything on the fused path is fp64 on fp32-stored bearings widened exactly, so blocks, cost and the
// LM trajectory agree with the fp64 oracle to ~1e-12; only the optional materialised Jacobian is fp32.
//
// Determinism: every reduction has a fixed order (lane-strided partial sums, xor-butterfly warp
// tree, ordered item sums per camera, fixed block tree) -- no floating-point atomics.
//
// One kernel launch per LM iteration: the evaluation kernel's last CTA to finish (atomic ticket)
// folds the work-item partials into per-camera blocks and, on a single GPU, also runs the
// trust-region decision + the damped 3x3 solves + the next candidate's rotation tables.  With an
// all-reduce installed (residual-sharded multi-GPU) the decision runs as its own 1-CTA kernel after
// the collective.
#include <cfloat>
#include <cmath>
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "bulk.cuh"
#include "geometry.cuh"

namespace sba {

struct CamParams {
    double Rd[9];    // d1 * R(r): X1 rotated = Rd . b1
    double dR[27];   // dR[k*9 + a*3 + b] = dR/dr_k [a][b]
    float nM[27];    // -d1 * dR/dr_k as fp32: the materialised Jacobian J[a][k] = sum_b nM[k][a][b] * b1[b]
    float pad;
};

struct Item {
    int64_t start;
    int count;
    int cam;
};

struct LMState {
    double cost, radius, dec_factor, model_dec, initial_cost;
    int iter, num_successful, termination, done;
    int consecutive_invalid, evals, max_iter, phase;
};

struct SolveConsts {
    double t[3];
    double d1, d2, huber;
};

constexpr int EVAL_THREADS = 256;
constexpr int EVAL_WARPS = EVAL_THREADS / 32;

// ---- rotation tables ----------------------------------------------------------------------------
// R(r) and dR/dr_k exactly as differentiating ceres::AngleAxisRotatePoint gives, both branches.
__device__ inline void rot_and_derivs(const double r[3], double R[9], double dR[3][9])
{
    double theta2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
    if (theta2 > DBL_EPSILON) {
        // one rsqrt instead of a square root and twelve divisions: this runs on the serial path of every LM iteration
        const double inv_th = rsqrt(theta2), th = theta2 * inv_th;
        double s, c;
        sincos(th, &s, &c);
        double w[3] = {r[0] * inv_th, r[1] * inv_th, r[2] * inv_th};
        double K[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) R[3 * a + b] = (a == b ? c : 0.0) + s * K[3 * a + b] + (1.0 - c) * w[a] * w[b];
        for (int k = 0; k < 3; k++) {
            double dw[3];
            for (int a = 0; a < 3; a++) dw[a] = ((a == k ? 1.0 : 0.0) - w[a] * w[k]) * inv_th;
            double dK[9] = {0, -dw[2], dw[1], dw[2], 0, -dw[0], -dw[1], dw[0], 0};
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++)
                    dR[k][3 * a + b] = (a == b ? -s * w[k] : 0.0) + c * w[k] * K[3 * a + b] + s * dK[3 * a + b] +
                                       s * w[k] * w[a] * w[b] + (1.0 - c) * (dw[a] * w[b] + w[a] * dw[b]);
        }
    } else {
        double K[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
        for (int a = 0; a < 9; a++) R[a] = K[a];
        R[0] += 1.0; R[4] += 1.0; R[8] += 1.0;
        for (int k = 0; k < 3; k++) {
            double e[3] = {k == 0 ? 1.0 : 0.0, k == 1 ? 1.0 : 0.0, k == 2 ? 1.0 : 0.0};
            double G[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
            for (int a = 0; a < 9; a++) dR[k][a] = G[a];
        }
    }
}

__device__ inline void write_cam_params(const double r[3], double d1, CamParams* out)
{
    double R[9], dR[3][9];
    rot_and_derivs(r, R, dR);
    for (int a = 0; a < 9; a++) out->Rd[a] = d1 * R[a];
    for (int k = 0; k < 3; k++)
        for (int a = 0; a < 9; a++) {
            out->dR[k * 9 + a] = dR[k][a];
            out->nM[k * 9 + a] = (float)(-d1 * dR[k][a]);
        }
    out->pad = 0.f;
}

__global__ void ba_cam_params_kernel(const double* __restrict__ x, int n_cam, double d1, CamParams* __restrict__ out)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cam) {
        double r[3] = {x[3 * c], x[3 * c + 1], x[3 * c + 2]};
        write_cam_params(r, d1, out + c);
    }
}

// Moments of one camera -> its normal-equation block.  m = {Sxx,Sxy,Sxz,Syy,Syz,Szz, C[b][a] (9), cost}
// with S = sum w b b^T, C[b][a] = sum w b_b res_a on UNSCALED bearings b (X1 = d1 b).
//   H_kl =  d1^2 sum_a sum_b sum_c dR_k[a][b] dR_l[a][c] S_bc        (J = -d1 dR b; the signs cancel)
//   g_k  = -d1   sum_a sum_b dR_k[a][b] C[b][a]
__device__ inline void moments_to_block(const double m[16], const CamParams* __restrict__ P, double d1, double blk[10])
{
    double dR[27];
#pragma unroll
    for (int i = 0; i < 27; i++) dR[i] = P->dR[i];
    const double S[9] = {m[0], m[1], m[2], m[1], m[3], m[4], m[2], m[4], m[5]};
    double T[27];   // T[k*9 + a*3 + c] = sum_b dR_k[a][b] S_bc
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                T[k * 9 + 3 * a + c] = dR[k * 9 + 3 * a] * S[c] + dR[k * 9 + 3 * a + 1] * S[3 + c] + dR[k * 9 + 3 * a + 2] * S[6 + c];
    int e = 0;
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
        for (int l = k; l < 3; l++) {
            double h = 0;
#pragma unroll
            for (int ac = 0; ac < 9; ac++) h += T[k * 9 + ac] * dR[l * 9 + ac];
            blk[e++] = d1 * d1 * h;
        }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double gk = 0;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++) gk += dR[k * 9 + 3 * a + b] * m[6 + 3 * b + a];
        blk[6 + k] = -d1 * gk;
    }
    blk[9] = m[15];
}

// Translation-only block (spherical_bundle_adjuster.cpp:948-1002): d res / d t = +I, so
// H = (sum w) I and g = sum w res.  m = {sum w, sum w res (3), ..., cost at [15]}.
__device__ inline void moments_to_block_tran(const double m[16], double blk[10])
{
    blk[0] = m[0]; blk[1] = 0; blk[2] = 0; blk[3] = m[0]; blk[4] = 0; blk[5] = m[0];
    blk[6] = m[1]; blk[7] = m[2]; blk[8] = m[3];
    blk[9] = m[15];
}

// ---- reductions ---------------------------------------------------------------------------------
__device__ inline double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Fixed-tree block sum; every thread gets the result.  `sh` has EVAL_THREADS doubles.
__device__ inline double block_sum(double v, double* sh)
{
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    double r = sh[0];
    __syncthreads();
    return r;
}

__device__ inline double block_max(double v, double* sh)
{
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = EVAL_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    double r = sh[0];
    __syncthreads();
    return r;
}

// ---- peer-memory exchange of the per-camera blocks (one CTA) ------------------------------------------------
constexpr int COMM_MAX_WORLD = 16;
// Peer waits give up after ~2 minutes of SM clock: long enough for ordinary rank skew (first-call plan builds, a rank busy on the
// host, a time-sliced GPU), short enough that a dead peer surfaces as a CUDA error instead of a hung box.
constexpr long long COMM_TIMEOUT_CYCLES = 240000000000ll;

constexpr int COMM_MAX_SLICES = 64;
constexpr int COMM_HEADER_BYTES = COMM_MAX_SLICES * 8;   // one sequence flag per slice, then the data

struct CommDev {
    double* data[COMM_MAX_WORLD];               // every rank's exchange buffer: [2 slots][stride] doubles
    volatile unsigned long long* flag[COMM_MAX_WORLD];   // every rank's per-slice sequence flags [COMM_MAX_SLICES]
    unsigned long long* seq;                    // this rank's private count of completed exchanges (local memory)
    unsigned int* ticket;                       // multi-CTA exchange: last CTA to finish advances seq
    int rank, world, stride;
};

// blk[0..count) <- sum over ranks, in rank order.  Called by ONE CTA with all EVAL_THREADS threads.
// Slot = sequence parity: a rank can only overwrite a slot two exchanges later, and to get there it must have
// seen every peer's NEXT flag, which a peer raises only after it finished reading this one.
__device__ void comm_exchange(const CommDev& cd, double* __restrict__ blk, int count)
{
    const unsigned long long seq = *cd.seq;
    const int slot = (int)(seq & 1) * cd.stride;
    double* mine = cd.data[cd.rank] + slot;
    for (int i = threadIdx.x; i < count; i += EVAL_THREADS) mine[i] = blk[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) cd.flag[cd.rank][0] = seq + 1;        // publish (slice 0 covers the whole vector here)
    if ((int)threadIdx.x < cd.world) {                            // one thread per peer waits for its flag
        const long long t0 = clock64();
        while (cd.flag[threadIdx.x][0] < seq + 1) {
            if (clock64() - t0 > COMM_TIMEOUT_CYCLES) __trap();  // a lost peer traps instead of hanging the GPU
        }
    }
    __threadfence_system();
    __syncthreads();
    for (int i = threadIdx.x; i < count; i += EVAL_THREADS) {
        double s = 0;
        for (int r = 0; r < cd.world; r++) s += __ldcv(cd.data[r] + slot + i);   // no stale cached copy of peer memory
        blk[i] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *cd.seq = seq + 1;
    __syncthreads();
}

// Stand-alone exchange for large block vectors (many cameras): CTA b owns slice b of the vector and its own
// sequence flag, so the NVLink reads of all slices run concurrently; the last CTA to finish advances seq.
__global__ void __launch_bounds__(EVAL_THREADS) ba_exchange_kernel(CommDev cd, double* blk, int count, int slice_len,
                                                                  const int* __restrict__ done)
{
    if (done && *done) return;
    const unsigned long long seq = *cd.seq;
    const int slot = (int)(seq & 1) * cd.stride;
    const int i0 = blockIdx.x * slice_len, i1 = min(count, i0 + slice_len);
    double* mine = cd.data[cd.rank] + slot;
    for (int i = i0 + threadIdx.x; i < i1; i += EVAL_THREADS) mine[i] = blk[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) cd.flag[cd.rank][blockIdx.x] = seq + 1;
    if ((int)threadIdx.x < cd.world) {
        const long long t0 = clock64();
        while (cd.flag[threadIdx.x][blockIdx.x] < seq + 1) {
            if (clock64() - t0 > COMM_TIMEOUT_CYCLES) __trap();
        }
    }
    __threadfence_system();
    __syncthreads();
    for (int i = i0 + threadIdx.x; i < i1; i += EVAL_THREADS) {
        double v[COMM_MAX_WORLD];
#pragma unroll
        for (int r = 0; r < COMM_MAX_WORLD; r++)
            if (r < cd.world) v[r] = __ldcv(cd.data[r] + slot + i);      // all peers' loads in flight together
        double sum = 0;
#pragma unroll
        for (int r = 0; r < COMM_MAX_WORLD; r++)
            if (r < cd.world) sum += v[r];                               // rank order
        blk[i] = sum;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(cd.ticket, 1u);
        if (tk == gridDim.x - 1) {
            *cd.ticket = 0;
            *cd.seq = seq + 1;
        }
    }
}

// ---- Ceres-default Levenberg-Marquardt decision (single CTA) ---------------------------------------
__device__ inline int solve3_spd(const double H[6], const double dd[3], const double rhs[3], double x[3])
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

struct LMArrays {
    double* x;         // n_cam x 3 current parameters
    double* xc;        // n_cam x 3 candidate
    double* blk_cur;   // n_cam x 10 blocks at x
    double* blk_cand;  // n_cam x 10 blocks at xc (just evaluated)
    double* scale;     // n_cam x 3 Jacobi scaling
    CamParams* params; // tables for xc
    LMState* st;
    int n_cam;
    double d1;         // uniform depth of camera 1 (folded into the rotation tables)
    int tran;          // 1: the free block is the translation (tables depend on the FIXED rotation: never rebuilt)
};

// Trust-region bookkeeping restated from Ceres' TrustRegionMinimizer / LevenbergMarquardtStrategy
// defaults (the reference only sets max_num_iterations=50, spherical_bundle_adjuster.cpp:334-338).
__device__ void lm_decide(const LMArrays A, double* sh)
{
    const double min_diag = 1e-6, max_diag = 1e32, min_rel_dec = 1e-3;
    const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, max_radius = 1e16, min_radius = 1e-32;
    __shared__ LMState S;
    __shared__ int s_accept;
    const int tid = threadIdx.x, nthr = EVAL_THREADS, n_cam = A.n_cam;

    if (tid == 0) S = *A.st;
    double part = 0;
    for (int c = tid; c < n_cam; c += nthr) part += A.blk_cand[10 * c + 9];
    double total_new = block_sum(part, sh);

    if (tid == 0) {
        s_accept = 0;
        S.evals++;
        if (S.phase == 0) {
            S.cost = total_new;
            S.initial_cost = total_new;
            S.phase = 1;
            s_accept = 2;  // adopt blocks, compute scaling
        } else {
            double cost_change = S.cost - total_new;
            if (fabs(cost_change) <= ftol * S.cost) {
                S.termination = 1; S.done = 1;
            } else {
                double rel = cost_change / S.model_dec;
                if (rel > min_rel_dec) {
                    s_accept = 1;
                    S.cost = total_new;
                    S.num_successful++;
                    double q = 2.0 * rel - 1.0;
                    S.radius = S.radius / fmax(1.0 / 3.0, 1.0 - q * q * q);
                    S.radius = fmin(max_radius, S.radius);
                    S.dec_factor = 2.0;
                } else {
                    S.radius = S.radius / S.dec_factor;
                    S.dec_factor *= 2.0;
                }
            }
        }
    }
    __syncthreads();
    const int accept = s_accept;
    if (accept) {
        double gm = 0;
        for (int c = tid; c < n_cam; c += nthr) {
            for (int k = 0; k < 10; k++) A.blk_cur[10 * c + k] = A.blk_cand[10 * c + k];
            for (int k = 0; k < 3; k++) {
                if (accept == 1) A.x[3 * c + k] = A.xc[3 * c + k];
                gm = fmax(gm, fabs(A.blk_cand[10 * c + 6 + k]));
            }
            if (accept == 2) {
                A.scale[3 * c] = 1.0 / (1.0 + sqrt(A.blk_cand[10 * c]));
                A.scale[3 * c + 1] = 1.0 / (1.0 + sqrt(A.blk_cand[10 * c + 3]));
                A.scale[3 * c + 2] = 1.0 / (1.0 + sqrt(A.blk_cand[10 * c + 5]));
            }
        }
        double gmax = block_max(gm, sh);
        if (tid == 0 && gmax <= gtol) { S.termination = 2; S.done = 1; }
    }
    __syncthreads();
    if (tid == 0 && !S.done && S.radius < min_radius) { S.termination = 5; S.done = 1; }   // Ceres: CONVERGENCE, "minimum trust region radius reached"
    __syncthreads();

    // next candidate (invalid steps retry in place: H and g do not change)
    while (!S.done) {
        __syncthreads();
        if (S.iter >= S.max_iter) {
            __syncthreads();
            if (tid == 0) { S.termination = 0; S.done = 1; }
            __syncthreads();
            break;
        }
        const double radius = S.radius;
        double md = 0, sn2 = 0, xn2 = 0, bad = 0;
        for (int c = tid; c < n_cam; c += nthr) {
            const double* B = A.blk_cur + 10 * c;
            const double* s = A.scale + 3 * c;
            double Hs[6] = {B[0] * s[0] * s[0], B[1] * s[0] * s[1], B[2] * s[0] * s[2],
                            B[3] * s[1] * s[1], B[4] * s[1] * s[2], B[5] * s[2] * s[2]};
            double gs[3] = {B[6] * s[0], B[7] * s[1], B[8] * s[2]};
            double dd[3] = {fmin(fmax(Hs[0], min_diag), max_diag) / radius, fmin(fmax(Hs[3], min_diag), max_diag) / radius,
                            fmin(fmax(Hs[5], min_diag), max_diag) / radius};
            double rhs[3] = {-gs[0], -gs[1], -gs[2]}, ds[3] = {0, 0, 0};
            if (solve3_spd(Hs, dd, rhs, ds)) bad = 1;
            double Hd[3] = {Hs[0] * ds[0] + Hs[1] * ds[1] + Hs[2] * ds[2], Hs[1] * ds[0] + Hs[3] * ds[1] + Hs[4] * ds[2],
                            Hs[2] * ds[0] + Hs[4] * ds[1] + Hs[5] * ds[2]};
            md -= (gs[0] * ds[0] + gs[1] * ds[1] + gs[2] * ds[2]) + 0.5 * (ds[0] * Hd[0] + ds[1] * Hd[1] + ds[2] * Hd[2]);
            for (int a = 0; a < 3; a++) {
                double st = ds[a] * s[a], xv = A.x[3 * c + a];
                A.xc[3 * c + a] = xv + st;
                sn2 += st * st;
                xn2 += xv * xv;
            }
        }
        md = block_sum(md, sh);
        sn2 = block_sum(sn2, sh);
        xn2 = block_sum(xn2, sh);
        bad = block_max(bad, sh);
        if (tid == 0) {
            S.iter++;
            if (bad > 0 || !(md > 0.0)) {
                if (++S.consecutive_invalid >= 5) { S.termination = 4; S.done = 1; }
                else {
                    S.radius *= 0.5;
                    if (S.radius < min_radius) { S.termination = 5; S.done = 1; }
                }
                s_accept = -1;
            } else {
                S.consecutive_invalid = 0;
                S.model_dec = md;
                // Ceres tests the parameter tolerance after evaluating the candidate; the test does
                // not depend on that evaluation, so it is applied here and the evaluation is saved.
                if (sqrt(sn2) <= ptol * (sqrt(xn2) + ptol)) { S.termination = 3; S.done = 1; }
                s_accept = 0;
            }
        }
        __syncthreads();
        if (s_accept == 0) break;
    }
    __syncthreads();
    if (!S.done && !A.tran)
        for (int c = tid; c < n_cam; c += nthr) {
            double r[3] = {A.xc[3 * c], A.xc[3 * c + 1], A.xc[3 * c + 2]};
            write_cam_params(r, A.d1, A.params + c);
        }
    if (tid == 0) *A.st = S;
}

__global__ void __launch_bounds__(EVAL_THREADS) ba_decide_kernel(LMArrays A)
{
    __shared__ double sh[EVAL_THREADS];
    if (A.st->done) return;
    lm_decide(A, sh);
}

// ---- the evaluation kernel (K5) ------------------------------------------------------------------
// Fold work-item moment partials into per-camera blocks, in item order.  `item_ptr` [n_cam+1] or NULL.
constexpr int NMOM = 16;

__device__ void fold_items(const double* __restrict__ partial, const int* __restrict__ item_ptr, int n_items, int n_cam,
                           const CamParams* __restrict__ params, double d1, int tran, double* __restrict__ blk)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (item_ptr == nullptr) {
        // single camera whose items are 0..n_items-1 (uniform layout): all 8 warps cooperate;
        // thread-strided ordered sums, then a fixed tree over the CTA
        __shared__ double red[EVAL_WARPS][NMOM];
        double acc[NMOM];
#pragma unroll
        for (int k = 0; k < NMOM; k++) acc[k] = 0;
        for (int it = threadIdx.x; it < n_items; it += EVAL_THREADS)
#pragma unroll
            for (int k = 0; k < NMOM; k++) acc[k] += partial[(size_t)it * NMOM + k];
#pragma unroll
        for (int k = 0; k < NMOM; k++) acc[k] = warp_sum(acc[k]);
        __syncthreads();
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < NMOM; k++) red[warp][k] = acc[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            double m[NMOM];
            for (int k = 0; k < NMOM; k++) {
                double t = 0;
                for (int w = 0; w < EVAL_WARPS; w++) t += red[w][k];
                m[k] = t;
            }
            if (tran) moments_to_block_tran(m, blk); else moments_to_block(m, params, d1, blk);
        }
        __syncthreads();
        return;
    }
    const int n_items_total = item_ptr[n_cam];
    if (n_items_total <= 8 * n_cam) {
        // few items per camera: one thread per camera, sequential ordered sum
        for (int c = threadIdx.x; c < n_cam; c += EVAL_THREADS) {
            double acc[NMOM];
#pragma unroll
            for (int k = 0; k < NMOM; k++) acc[k] = 0;
            for (int it = item_ptr[c]; it < item_ptr[c + 1]; it++)
#pragma unroll
                for (int k = 0; k < NMOM; k++) acc[k] += partial[(size_t)it * NMOM + k];
            if (tran) moments_to_block_tran(acc, blk + 10 * c); else moments_to_block(acc, params + c, d1, blk + 10 * c);
        }
    } else {
        // many items per camera: one warp per camera, lane-strided ordered sums + butterfly
        for (int c = warp; c < n_cam; c += EVAL_WARPS) {
            double acc[NMOM];
#pragma unroll
            for (int k = 0; k < NMOM; k++) acc[k] = 0;
            for (int it = item_ptr[c] + lane; it < item_ptr[c + 1]; it += 32)
#pragma unroll
                for (int k = 0; k < NMOM; k++) acc[k] += partial[(size_t)it * NMOM + k];
#pragma unroll
            for (int k = 0; k < NMOM; k++) acc[k] = warp_sum(acc[k]);
            if (lane == 0) { if (tran) moments_to_block_tran(acc, blk + 10 * c); else moments_to_block(acc, params + c, d1, blk + 10 * c); }
        }
    }
}

// Stand-alone fold for problems with many cameras: one warp per camera over the whole grid (the
// in-kernel fold runs in a single CTA, which is only right when there are few cameras).
__global__ void __launch_bounds__(EVAL_THREADS) ba_fold_kernel(const double* __restrict__ partial, const int* __restrict__ item_ptr, int n_cam,
                                                              const CamParams* __restrict__ params, double d1, int tran,
                                                              double* __restrict__ blk, const int* __restrict__ done)
{
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * EVAL_WARPS + (threadIdx.x >> 5);
    if (c >= n_cam || (done && *done)) return;
    double acc[NMOM];
#pragma unroll
    for (int k = 0; k < NMOM; k++) acc[k] = 0;
    for (int it = item_ptr[c] + lane; it < item_ptr[c + 1]; it += 32)
#pragma unroll
        for (int k = 0; k < NMOM; k++) acc[k] += partial[(size_t)it * NMOM + k];
#pragma unroll
    for (int k = 0; k < NMOM; k++) acc[k] = warp_sum(acc[k]);
    if (lane == 0) { if (tran) moments_to_block_tran(acc, blk + 10 * c); else moments_to_block(acc, params + c, d1, blk + 10 * c); }
}

struct EvalArgs {
    const float4* b1;
    const float4* b2;
    const int32_t* perm;  // sorted position -> caller's observation index (NULL = identity)
    const Item* items;     // NULL = uniform layout: one camera, item i covers [i*item_len, min(n_obs, (i+1)*item_len))
    const int* item_ptr;   // NULL with the uniform layout
    int item_len;
    int n_obs;              // observation count, or the CAPACITY when n_obs_dev is given
    const int* n_obs_dev;   // optional: the actual count lives on the device (fused pair pipeline)
    int n_items;
    int n_cam;
    const CamParams* params;
    double* partial;
    double* blk_out;
    unsigned int* ticket;
    const double* tvec;  // translation-only mode: per-camera t [n_cam x 3] on the device (NULL: the uniform k.t)
    int tran;            // 1: accumulate the translation-block moments {sum w, sum w res}
    const CommDev* comm;  // peer-memory exchange after the fold (NULL: single GPU or host all-reduce callback)
    const int* done;  // LM solve: skip the whole evaluation once the solver has converged (NULL = always run)
    float* res;  // optional materialised outputs (caller order)
    float* jac;
    SolveConsts k;
};

// ---- observation staging: per-warp shared-memory ring filled by TMA 1-D bulk copies ----------------------
// Every warp streams its work items through BULK_STAGES x 32 observations of shared memory.  One lane
// issues `cp.async.bulk` for the next chunk of b1 and b2 (512 B each) and the bytes land asynchronously
// on a per-warp mbarrier, so the memory-level parallelism (3 chunks = 3 KB in flight per warp, 72 KB per
// SM at 24 resident warps) is fixed by construction instead of by register allocation and instruction
// scheduling of explicit prefetch loads.
constexpr int BULK_STAGES = 4;

// MODE 0: partials only (ba_fold_kernel follows); 1: the last CTA folds; 2: the last CTA folds and runs the LM decision.
// DEVN: the observation count is read from device memory (fused pair pipeline); a separate instantiation
// because routing the common path's count through that load measurably de-tunes its inner loop.
template <bool WRITE, int MODE, bool TRAN, bool DEVN>
__global__ void __launch_bounds__(EVAL_THREADS, 3) ba_rot_eval_kernel(EvalArgs E, LMArrays A)
{
    __shared__ double sh[EVAL_THREADS];
    __shared__ int s_last;
    // per-warp copy of the current camera's tables: d1*R (9 fp64) and, for the materialised variant,
    // -d1*dR/dr_k (27 fp32).  Shared memory (broadcast reads) instead of registers: three CTAs per SM.
    __shared__ __align__(16) double s_R[EVAL_WARPS][10];
    __shared__ __align__(16) float s_M[WRITE ? EVAL_WARPS : 1][28];
    __shared__ __align__(128) float4 s_obs[EVAL_WARPS][BULK_STAGES][2][32];   // [warp][stage][b1|b2][lane]
    __shared__ __align__(8) uint64_t s_bar[EVAL_WARPS][BULK_STAGES];
    if (E.done && *E.done) return;  // converged earlier in this launch chunk

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp_global = blockIdx.x * EVAL_WARPS + wib;
    const int warp_stride = gridDim.x * EVAL_WARPS;
    const double d2 = E.k.d2, huber = E.k.huber, hub2 = huber * huber;
    const double t0 = E.k.t[0], t1 = E.k.t[1], t2 = E.k.t[2];
    const double* Rs = &s_R[wib][0];
    const int n_obs = DEVN ? min(E.n_obs, *E.n_obs_dev) : E.n_obs;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < BULK_STAGES; k++)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&s_bar[wib][k])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t gchunk = 0;   // chunks this warp has consumed so far: ring position and mbarrier phase
    // Uniform layout (one camera): a warp keeps its moments across all of its work items and the CTA
    // publishes ONE partial, so the final fold is over gridDim.x entries instead of n_items.
    const bool uniform = (E.items == nullptr);
    double acc[NMOM];
#pragma unroll
    for (int k = 0; k < NMOM; k++) acc[k] = 0;

    for (int it = warp_global; it < E.n_items; it += warp_stride) {
        Item item;
        if (E.items) item = E.items[it];
        else {
            item.start = (int64_t)it * E.item_len;
            item.count = min(E.item_len, n_obs - it * E.item_len);
            item.cam = 0;
        }
        // chunks of 32 observations; the first BULK_STAGES-1 go in flight before the tables are staged
        const float4* g1 = E.b1 + item.start;
        const float4* g2 = E.b2 + item.start;
        const int n_chunks = (item.count + 31) >> 5;
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < BULK_STAGES - 1; c++)
                if (c < n_chunks) {
                    const int st = (int)((gchunk + c) % BULK_STAGES);
                    const uint32_t bytes = (uint32_t)min(32, item.count - 32 * c) * 16u;
                    bar_expect(&s_bar[wib][st], 2 * bytes);
                    bulk_load(&s_obs[wib][st][0][0], g1 + 32 * c, bytes, &s_bar[wib][st]);
                    bulk_load(&s_obs[wib][st][1][0], g2 + 32 * c, bytes, &s_bar[wib][st]);
                }
        }
        {
            const CamParams* P = E.params + item.cam;
            __syncwarp();
            if (lane < 9) s_R[wib][lane] = P->Rd[lane];
            if (WRITE && lane < 27) s_M[WRITE ? wib : 0][lane] = P->nM[lane];
            __syncwarp();
        }
        // translation-only mode: this camera's candidate t replaces the uniform one
        const double u0 = TRAN ? E.tvec[3 * item.cam] : t0, u1 = TRAN ? E.tvec[3 * item.cam + 1] : t1, u2 = TRAN ? E.tvec[3 * item.cam + 2] : t2;

        if (!uniform) {
#pragma unroll
            for (int k = 0; k < NMOM; k++) acc[k] = 0;
        }

        for (int c = 0; c < n_chunks; c++, gchunk++) {
            // refill the stage that was consumed one iteration ago (the __syncwarp below ordered its reads)
            if (lane == 0 && c + BULK_STAGES - 1 < n_chunks) {
                const int cc = c + BULK_STAGES - 1;
                const int st = (int)((gchunk + BULK_STAGES - 1) % BULK_STAGES);
                const uint32_t bytes = (uint32_t)min(32, item.count - 32 * cc) * 16u;
                bar_expect(&s_bar[wib][st], 2 * bytes);
                bulk_load(&s_obs[wib][st][0][0], g1 + 32 * cc, bytes, &s_bar[wib][st]);
                bulk_load(&s_obs[wib][st][1][0], g2 + 32 * cc, bytes, &s_bar[wib][st]);
            }
            const int st = (int)(gchunk % BULK_STAGES);
            bar_wait(&s_bar[wib][st], (uint32_t)((gchunk / BULK_STAGES) & 1));
            const int o = 32 * c + lane;
            if (o < item.count) {
            const float4 p1 = s_obs[wib][st][0][lane], p2 = s_obs[wib][st][1][lane];
            // residual in fp64: res = d2*b2 - (d1*R*b1 - t)   (spherical_bundle_adjuster.cpp:896-916)
            const double bx = (double)p1.x, by = (double)p1.y, bz = (double)p1.z;
            const double rx = fma(d2, (double)p2.x, u0) - (Rs[0] * bx + Rs[1] * by + Rs[2] * bz);
            const double ry = fma(d2, (double)p2.y, u1) - (Rs[3] * bx + Rs[4] * by + Rs[5] * bz);
            const double rz = fma(d2, (double)p2.z, u2) - (Rs[6] * bx + Rs[7] * by + Rs[8] * bz);
            const double s = rx * rx + ry * ry + rz * rz;
            // Huber: rho' = 1 (s <= a^2) or a/sqrt(s); rho = s or 2 a sqrt(s) - a^2
            double rho = s, w = 1.0;
            if (huber > 0.0 && s > hub2) {
                float y0;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)s));   // seed, refined to fp64 below
                double y = (double)y0;
                y = y * (1.5 - 0.5 * s * y * y);
                y = y * (1.5 - 0.5 * s * y * y);
                w = huber * y;
                rho = 2.0 * huber * (s * y) - hub2;
            }
            // weighted moments: S += w b b^T (6), C[b][a] += w b_b res_a (9), cost += rho/2
            if (TRAN) {   // translation block: J = +I
                acc[0] += w;
                acc[1] = fma(w, rx, acc[1]); acc[2] = fma(w, ry, acc[2]); acc[3] = fma(w, rz, acc[3]);
                acc[15] = fma(0.5, rho, acc[15]);
                if (WRITE && E.res) {
                    const int64_t i = item.start + o;
                    const int64_t dst = E.perm ? (int64_t)E.perm[i] : i;
                    E.res[3 * dst] = (float)rx; E.res[3 * dst + 1] = (float)ry; E.res[3 * dst + 2] = (float)rz;
                }
            } else {
            const double wx = w * bx, wy = w * by, wz = w * bz;
            acc[0] = fma(wx, bx, acc[0]); acc[1] = fma(wx, by, acc[1]); acc[2] = fma(wx, bz, acc[2]);
            acc[3] = fma(wy, by, acc[3]); acc[4] = fma(wy, bz, acc[4]); acc[5] = fma(wz, bz, acc[5]);
            acc[6] = fma(wx, rx, acc[6]); acc[7] = fma(wx, ry, acc[7]); acc[8] = fma(wx, rz, acc[8]);
            acc[9] = fma(wy, rx, acc[9]); acc[10] = fma(wy, ry, acc[10]); acc[11] = fma(wy, rz, acc[11]);
            acc[12] = fma(wz, rx, acc[12]); acc[13] = fma(wz, ry, acc[13]); acc[14] = fma(wz, rz, acc[14]);
            acc[15] = fma(0.5, rho, acc[15]);
            if (WRITE) {
                // materialised RAW residual and Jacobian (fp32), in the caller's observation order
                const int64_t i = item.start + o;
                const int64_t dst = E.perm ? (int64_t)E.perm[i] : i;
                if (E.res) {
                    E.res[3 * dst] = (float)rx; E.res[3 * dst + 1] = (float)ry; E.res[3 * dst + 2] = (float)rz;
                }
                if (E.jac) {
                    const float* Ms = &s_M[WRITE ? wib : 0][0];
                    float J[9];
#pragma unroll
                    for (int k = 0; k < 3; k++)
#pragma unroll
                        for (int a = 0; a < 3; a++) J[3 * a + k] = Ms[k * 9 + a * 3] * p1.x + Ms[k * 9 + a * 3 + 1] * p1.y + Ms[k * 9 + a * 3 + 2] * p1.z;
#pragma unroll
                    for (int a = 0; a < 9; a++) E.jac[9 * dst + a] = J[a];
                }
            }
            }   // !TRAN
            }   // o < item.count
            __syncwarp();   // every lane is done with this stage before lane 0 refills it
        }
        if (!uniform) {
#pragma unroll
            for (int k = 0; k < NMOM; k++) acc[k] = warp_sum(acc[k]);
            if (lane == 0) {
                double* out = E.partial + (size_t)it * NMOM;
#pragma unroll
                for (int k = 0; k < NMOM; k++) out[k] = acc[k];
            }
        }
    }
    if (uniform) {
        // CTA partial: butterfly inside each warp, then the 8 warps in order
        __shared__ double s_cta[EVAL_WARPS][NMOM];
#pragma unroll
        for (int k = 0; k < NMOM; k++) acc[k] = warp_sum(acc[k]);
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < NMOM; k++) s_cta[wib][k] = acc[k];
        __syncthreads();
        if (threadIdx.x < NMOM) {
            double t = 0;
            for (int w = 0; w < EVAL_WARPS; w++) t += s_cta[w][threadIdx.x];
            E.partial[(size_t)blockIdx.x * NMOM + threadIdx.x] = t;
        }
    }

    if (MODE == 0) return;
    // last CTA to finish folds the partials (classic threadfence reduction ticket)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int tk = atomicAdd(E.ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    fold_items(E.partial, E.item_ptr, E.items ? E.n_items : (int)gridDim.x, E.n_cam, E.params, E.k.d1, TRAN ? 1 : 0, E.blk_out);
    if (threadIdx.x == 0) *E.ticket = 0;
    if (E.comm) {          // sharded residuals: sum every rank's blocks over NVLink before anything looks at them
        __syncthreads();
        comm_exchange(*E.comm, E.blk_out, E.n_cam * 10);
    }
    if (MODE == 2) {
        __syncthreads();
        lm_decide(A, sh);
    }
}

// ---- one launch for a whole image pair's rotation solve (fused pair pipeline) ------------------------------------------
// A C2-sized pair has ~8 k matches: one evaluation is 256 KB of bearings and under a microsecond of arithmetic spread
// over a few SMs, so a solve made of one launch per LM evaluation (plus the bearing and table kernels before it) is
// launch gaps and global-memory round trips, not work.  This kernel does everything between the match list and the
// rotation in ONE launch of ONE thread-block cluster (16 CTAs, falling back to 8):
//   1. matched keypoints -> ERP pixels -> unit bearings (equi2cube_surf.cpp:96-113 gather + cube2equi_pixel,
//      spherical_bundle_adjuster.cpp:271-298), written to b1/b2 (each thread later re-reads only what it wrote);
//   2. the LM loop: every CTA accumulates the moments of its observations (same arithmetic as ba_rot_eval_kernel)
//      and stores its 16 sums straight into CTA 0's shared memory (distributed shared memory); after a cluster
//      barrier CTA 0 adds them in rank order, runs the trust-region decision (lm_decide_single: lm_decide with the
//      one-camera sums written out, one thread) and pushes the next candidate's rotation table and the done flag into
//      every CTA's shared memory; second cluster barrier.  No global-memory hand-over inside an iteration.
constexpr int PAIR_SOLVE_CLUSTER = 16;
constexpr int PAIR_SOLVE_THREADS = 512;
constexpr int PAIR_SOLVE_WARPS = PAIR_SOLVE_THREADS / 32;

struct PairSolveArgs {
    const float2* key_l;
    const float2* key_r;
    const int32_t* qi;
    const int32_t* ti;
    const int32_t* d_n;   // match count on the device
    int cap, cs, w, h;
    float4* b1;
    float4* b2;
    SolveConsts k;
    // starting point and fresh solver state travel as kernel arguments, the result goes straight into the problem's pinned host
    // mailboxes: no copy nodes on the stream in front of or behind the launch
    double r0[3];
    LMState init;
    LMState* h_state;
    double* h_x;
};

// Damped 3x3 solve for the one-thread decision of the fused pair kernel.  Same Cholesky as solve3_spd, but every
// division / square root is a reciprocal (__drcp_rn) or rsqrt() (inline MUFU seed + Newton, <= 1 ulp) times a multiply:
// a correctly rounded fp64 division or square root is a ~40-instruction subroutine, and this code runs once, serially,
// on the critical path of every LM iteration.  Results differ from solve3_spd in the last bit or two.
__device__ inline int solve3_spd_fast(const double H[6], const double dd[3], const double rhs[3], double x[3])
{
    const double a00 = H[0] + dd[0], a01 = H[1], a02 = H[2], a11 = H[3] + dd[1], a12 = H[4], a22 = H[5] + dd[2];
    if (!(a00 > 0.0)) return 1;
    const double i00 = rsqrt(a00), l10 = a01 * i00, l20 = a02 * i00;      // 1 / l00
    const double t11 = a11 - l10 * l10;
    if (!(t11 > 0.0)) return 1;
    const double i11 = rsqrt(t11), l21 = (a12 - l20 * l10) * i11;
    const double t22 = a22 - l20 * l20 - l21 * l21;
    if (!(t22 > 0.0)) return 1;
    const double i22 = rsqrt(t22);
    const double y0 = rhs[0] * i00, y1 = (rhs[1] - l10 * y0) * i11, y2 = (rhs[2] - l20 * y0 - l21 * y1) * i22;
    x[2] = y2 * i22;
    x[1] = (y1 - l21 * x[2]) * i11;
    x[0] = (y0 - l10 * x[1] - l20 * x[2]) * i00;
    return 0;
}

// lm_decide for ONE camera, one thread: every block_sum / block_max of lm_decide is over a single term here, so the
// decisions (and the bits) are the same.  Returns with *A.st updated; A.params holds the tables of the new candidate.
__device__ void lm_decide_single(const LMArrays A)
{
    const double min_diag = 1e-6, max_diag = 1e32, min_rel_dec = 1e-3;
    const double ftol = 1e-6, gtol = 1e-10, ptol = 1e-8, max_radius = 1e16, min_radius = 1e-32;
    LMState S = *A.st;
    const double total_new = A.blk_cand[9];
    int accept = 0;
    S.evals++;
    if (S.phase == 0) {
        S.cost = total_new; S.initial_cost = total_new; S.phase = 1;
        accept = 2;
    } else {
        const double cost_change = S.cost - total_new;
        if (fabs(cost_change) <= ftol * S.cost) { S.termination = 1; S.done = 1; }
        else {
            const double rel = cost_change * __drcp_rn(S.model_dec);
            if (rel > min_rel_dec) {
                accept = 1;
                S.cost = total_new;
                S.num_successful++;
                const double q = 2.0 * rel - 1.0;
                S.radius = S.radius * __drcp_rn(fmax(1.0 / 3.0, 1.0 - q * q * q));
                S.radius = fmin(max_radius, S.radius);
                S.dec_factor = 2.0;
            } else {
                S.radius = S.radius * __drcp_rn(S.dec_factor);   // dec_factor is a power of two: exact
                S.dec_factor *= 2.0;
            }
        }
    }
    if (accept) {
        double gm = 0;
        for (int k = 0; k < 10; k++) A.blk_cur[k] = A.blk_cand[k];
        for (int k = 0; k < 3; k++) {
            if (accept == 1) A.x[k] = A.xc[k];
            gm = fmax(gm, fabs(A.blk_cand[6 + k]));
        }
        if (accept == 2) {
            // 1 / (1 + sqrt(h)) with sqrt(h) = h * rsqrt(h)
            const double h0 = A.blk_cand[0], h1 = A.blk_cand[3], h2 = A.blk_cand[5];
            A.scale[0] = __drcp_rn(1.0 + (h0 > 0.0 ? h0 * rsqrt(h0) : 0.0));
            A.scale[1] = __drcp_rn(1.0 + (h1 > 0.0 ? h1 * rsqrt(h1) : 0.0));
            A.scale[2] = __drcp_rn(1.0 + (h2 > 0.0 ? h2 * rsqrt(h2) : 0.0));
        }
        if (gm <= gtol) { S.termination = 2; S.done = 1; }
    }
    if (!S.done && S.radius < min_radius) { S.termination = 5; S.done = 1; }
    while (!S.done) {
        if (S.iter >= S.max_iter) { S.termination = 0; S.done = 1; break; }
        const double radius = S.radius;
        const double* B = A.blk_cur;
        const double* s = A.scale;
        double Hs[6] = {B[0] * s[0] * s[0], B[1] * s[0] * s[1], B[2] * s[0] * s[2], B[3] * s[1] * s[1], B[4] * s[1] * s[2], B[5] * s[2] * s[2]};
        double gs[3] = {B[6] * s[0], B[7] * s[1], B[8] * s[2]};
        const double inv_radius = __drcp_rn(radius);
        double dd[3] = {fmin(fmax(Hs[0], min_diag), max_diag) * inv_radius, fmin(fmax(Hs[3], min_diag), max_diag) * inv_radius,
                        fmin(fmax(Hs[5], min_diag), max_diag) * inv_radius};
        double rhs[3] = {-gs[0], -gs[1], -gs[2]}, ds[3] = {0, 0, 0};
        const int bad = solve3_spd_fast(Hs, dd, rhs, ds);
        double Hd[3] = {Hs[0] * ds[0] + Hs[1] * ds[1] + Hs[2] * ds[2], Hs[1] * ds[0] + Hs[3] * ds[1] + Hs[4] * ds[2],
                        Hs[2] * ds[0] + Hs[4] * ds[1] + Hs[5] * ds[2]};
        double md = 0, sn2 = 0, xn2 = 0;
        md -= (gs[0] * ds[0] + gs[1] * ds[1] + gs[2] * ds[2]) + 0.5 * (ds[0] * Hd[0] + ds[1] * Hd[1] + ds[2] * Hd[2]);
        for (int a = 0; a < 3; a++) {
            const double st = ds[a] * s[a], xv = A.x[a];
            A.xc[a] = xv + st;
            sn2 += st * st;
            xn2 += xv * xv;
        }
        S.iter++;
        if (bad || !(md > 0.0)) {
            if (++S.consecutive_invalid >= 5) { S.termination = 4; S.done = 1; }
            else {
                S.radius *= 0.5;
                if (S.radius < min_radius) { S.termination = 5; S.done = 1; }
            }
            continue;   // retry in place: H and g do not change
        }
        S.consecutive_invalid = 0;
        S.model_dec = md;
        if ((sn2 > 0.0 ? sn2 * rsqrt(sn2) : 0.0) <= ptol * ((xn2 > 0.0 ? xn2 * rsqrt(xn2) : 0.0) + ptol)) { S.termination = 3; S.done = 1; }
        break;
    }
    if (!S.done) {
        double r[3] = {A.xc[0], A.xc[1], A.xc[2]};
        write_cam_params(r, A.d1, A.params);
    }
    *A.st = S;
}

}  // namespace sba

#include <cooperative_groups.h>

namespace sba {

#ifdef SBA_TC_TRACE
// Debug build only (make trace): time stamps of rank 0 / thread 0, read back through sba_ps_trace_read().
__device__ unsigned long long g_ps_trace[64];
#define PS_TRACE(slot) do { if (rank == 0 && tid == 0 && (slot) < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_ps_trace[(slot)] = t_; } } while (0)
#else
#define PS_TRACE(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(PAIR_SOLVE_THREADS) ba_pair_solve_kernel(PairSolveArgs P, LMArrays A)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int rank = cluster.block_rank(), nblk = cluster.num_blocks();
    __shared__ double s_cta[PAIR_SOLVE_WARPS][NMOM];
    __shared__ double s_part[PAIR_SOLVE_CLUSTER][NMOM];   // rank 0's copy receives every CTA's sums
    __shared__ double s_R[9];
    __shared__ int s_done;
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    pdl_wait();   // launched as a programmatic dependent of knn2_finalize_kernel: the match list and its count are complete from here
    const int n = min(P.cap, *P.d_n);
    const int stride = nblk * PAIR_SOLVE_THREADS;
    const double d1 = P.k.d1, d2 = P.k.d2, huber = P.k.huber, hub2 = huber * huber;
    const double t0 = P.k.t[0], t1 = P.k.t[1], t2 = P.k.t[2];
    int gen_ = 0;
    PS_TRACE(0);
    // The starting rotation's tables.  Every CTA builds its own copy of d1 * R (thread 0, before its bearing), and a SECOND thread
    // of CTA 0 writes the derivative tables the first contraction needs -- nothing is broadcast and nobody waits for a single
    // thread's serial fp64 code here; the bearings below hide it.
    if (tid == 0) {
        double r[3] = {P.r0[0], P.r0[1], P.r0[2]}, R[9], dR[3][9];
        if (rank == 0) {   // the solver's device-side state starts here (the decision below is this same thread)
#pragma unroll
            for (int a = 0; a < 3; a++) { A.x[a] = r[a]; A.xc[a] = r[a]; }
            *A.st = P.init;
        }
        rot_and_derivs(r, R, dR);
#pragma unroll
        for (int a = 0; a < 9; a++) s_R[a] = d1 * R[a];
        s_done = 0;
    }
    if (rank == 0 && tid == 32) {
        double r[3] = {P.r0[0], P.r0[1], P.r0[2]};
        write_cam_params(r, d1, A.params);
    }
    // 1. bearings of this thread's matches
    for (int i = rank * PAIR_SOLVE_THREADS + tid; i < n; i += stride) {
        const float2 kl = P.key_l[P.qi[i]], kr = P.key_r[P.ti[i]];
        float ex, ey;
        double x, y, z;
        cube2equi_point(kl.x, kl.y, P.cs, P.w, P.h, &ex, &ey);
        pixel_to_bearing(ex, ey, (double)P.w, (double)P.h, &x, &y, &z);
        P.b1[i] = make_float4((float)x, (float)y, (float)z, 0.f);
        cube2equi_point(kr.x, kr.y, P.cs, P.w, P.h, &ex, &ey);
        pixel_to_bearing(ex, ey, (double)P.w, (double)P.h, &x, &y, &z);
        P.b2[i] = make_float4((float)x, (float)y, (float)z, 0.f);
    }
    PS_TRACE(1);
    cluster.sync();                // bearings done; every CTA of the cluster is running: shared memory may be written remotely from here on
    PS_TRACE(2);
    // 2. the LM loop: one pass per evaluation
    while (true) {
        double acc[NMOM];
#pragma unroll
        for (int k = 0; k < NMOM; k++) acc[k] = 0;
        for (int i = rank * PAIR_SOLVE_THREADS + tid; i < n; i += stride) {
            const float4 p1 = P.b1[i], p2 = P.b2[i];
            const double bx = (double)p1.x, by = (double)p1.y, bz = (double)p1.z;
            const double rx = fma(d2, (double)p2.x, t0) - (s_R[0] * bx + s_R[1] * by + s_R[2] * bz);
            const double ry = fma(d2, (double)p2.y, t1) - (s_R[3] * bx + s_R[4] * by + s_R[5] * bz);
            const double rz = fma(d2, (double)p2.z, t2) - (s_R[6] * bx + s_R[7] * by + s_R[8] * bz);
            const double s = rx * rx + ry * ry + rz * rz;
            double rho = s, w = 1.0;
            if (huber > 0.0 && s > hub2) {   // Huber: rho' = a / sqrt(s), rho = 2 a sqrt(s) - a^2 (same refinement as ba_rot_eval_kernel)
                float y0;
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)s));
                double y = (double)y0;
                y = y * (1.5 - 0.5 * s * y * y);
                y = y * (1.5 - 0.5 * s * y * y);
                w = huber * y;
                rho = 2.0 * huber * (s * y) - hub2;
            }
            const double wx = w * bx, wy = w * by, wz = w * bz;
            acc[0] = fma(wx, bx, acc[0]); acc[1] = fma(wx, by, acc[1]); acc[2] = fma(wx, bz, acc[2]);
            acc[3] = fma(wy, by, acc[3]); acc[4] = fma(wy, bz, acc[4]); acc[5] = fma(wz, bz, acc[5]);
            acc[6] = fma(wx, rx, acc[6]); acc[7] = fma(wx, ry, acc[7]); acc[8] = fma(wx, rz, acc[8]);
            acc[9] = fma(wy, rx, acc[9]); acc[10] = fma(wy, ry, acc[10]); acc[11] = fma(wy, rz, acc[11]);
            acc[12] = fma(wz, rx, acc[12]); acc[13] = fma(wz, ry, acc[13]); acc[14] = fma(wz, rz, acc[14]);
            acc[15] = fma(0.5, rho, acc[15]);
        }
#pragma unroll
        for (int k = 0; k < NMOM; k++) acc[k] = warp_sum(acc[k]);
        if (lane == 0)
#pragma unroll
            for (int k = 0; k < NMOM; k++) s_cta[wib][k] = acc[k];
        __syncthreads();
        if (tid < NMOM) {   // this CTA's sums -> CTA 0's shared memory
            double t = 0;
            for (int w8 = 0; w8 < PAIR_SOLVE_WARPS; w8++) t += s_cta[w8][tid];
            cluster.map_shared_rank(&s_part[0][0], 0)[rank * NMOM + tid] = t;
        }
        PS_TRACE(3 + 4 * gen_);
        cluster.sync();
        PS_TRACE(4 + 4 * gen_);
        if (rank == 0) {   // warp 0: fold in rank order (16 lanes), moments -> block and the decision (lane 0), broadcast (one lane per CTA)
            if (tid < NMOM) {
                double t = 0;
                for (unsigned int q = 0; q < nblk; q++) t += s_part[q][tid];   // rank order
                s_cta[0][tid] = t;
            }
            __syncwarp();
            if (tid == 0) {
                moments_to_block(&s_cta[0][0], A.params, d1, A.blk_cand);   // A.params: the tables of the candidate just evaluated
                lm_decide_single(A);                                         // ... and now of the next one
            }
            __syncwarp();
            if (tid < (int)nblk) {
                const int done = *((volatile int*)&A.st->done);
                double* dst = cluster.map_shared_rank(s_R, tid);
                if (!done)
#pragma unroll
                    for (int a = 0; a < 9; a++) dst[a] = ((volatile double*)A.params->Rd)[a];
                *cluster.map_shared_rank(&s_done, tid) = done;
            }
        }
        PS_TRACE(5 + 4 * gen_);
        cluster.sync();
        PS_TRACE(6 + 4 * gen_);
        gen_++;
        if (s_done) break;
    }
    if (rank == 0 && tid == 0) {   // result -> pinned host mailboxes (read by ba_solve_finish after its synchronise)
        static_assert(sizeof(LMState) % 8 == 0, "LMState is copied as 64-bit words");
#pragma unroll
        for (int a = 0; a < (int)(sizeof(LMState) / 8); a++) ((unsigned long long*)P.h_state)[a] = ((volatile unsigned long long*)A.st)[a];
#pragma unroll
        for (int a = 0; a < 3; a++) P.h_x[a] = ((volatile double*)A.x)[a];
        __threadfence_system();
    }
}

// blocks [n_cam x 10] -> H [n_cam x 6], g [n_cam x 3], cost [n_cam]
__global__ void ba_unpack_blocks_kernel(const double* __restrict__ blk, int n_cam, double* __restrict__ H, double* __restrict__ g,
                                        double* __restrict__ cost)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cam) return;
    if (H) for (int k = 0; k < 6; k++) H[6 * c + k] = blk[10 * c + k];
    if (g) for (int k = 0; k < 3; k++) g[3 * c + k] = blk[10 * c + 6 + k];
    if (cost) cost[c] = blk[10 * c + 9];
}

// ---- problem construction ---------------------------------------------------------------------------
__global__ void iota_kernel(int32_t* p, int64_t n)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (int32_t)i;
}

__global__ void cam_hist_kernel(const int32_t* __restrict__ cam, int64_t n, int n_cam, int* __restrict__ counts, int* __restrict__ bad)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = cam[i];
        if (c < 0 || c >= n_cam) atomicExch(bad, 1);
        else atomicAdd(counts + c, 1);
    }
}

__global__ void gather_obs_kernel(const float4* __restrict__ b1, const float4* __restrict__ b2, const int32_t* __restrict__ perm, int64_t n,
                                  float4* __restrict__ o1, float4* __restrict__ o2)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int32_t s = perm[i];
        o1[i] = b1[s];
        o2[i] = b2[s];
    }
}

}  // namespace sba

using namespace sba;

struct sba_ba_problem {
    sba_ctx* ctx = nullptr;
    int64_t n_obs = 0;
    int n_cam = 0, n_items = 0;
    float4* b1 = nullptr;
    float4* b2 = nullptr;
    int32_t* perm = nullptr;
    Item* items = nullptr;
    int* item_ptr = nullptr;
    double* partial = nullptr;
    CamParams* params = nullptr;
    double *x = nullptr, *xc = nullptr, *blk_cur = nullptr, *blk_cand = nullptr, *scale = nullptr;
    LMState* state = nullptr;
    unsigned int* ticket = nullptr;
    LMState* h_state = nullptr;  // pinned
    double* h_x = nullptr;       // pinned, n_cam x 3
    sba_allreduce_fn allreduce = nullptr;
    void* allreduce_user = nullptr;
    int eval_blocks = 1;
    int item_len = 32;
    const CommDev* comm_host = nullptr; // host copy of the descriptor (kernel argument of the stand-alone exchange)
    const CommDev* comm_dev = nullptr;  // device copy of the peer-exchange descriptor (owned by the sba_comm)
    double* fixed = nullptr;          // translation-only solves: the fixed rotations [n_cam x 3] on the device
    const int* n_obs_dev = nullptr;   // actual observation count on the device (n_obs is then the capacity)
    bool borrowed = false;   // b1/b2 belong to the caller (fused pipeline): not returned to the cache
};

static void free_problem(sba_ba_problem* p)
{
    if (!p) return;
    sba::BlockCache& C = p->ctx->cache;
    if (p->borrowed) p->b1 = p->b2 = nullptr;
    void* dev[] = {p->b1, p->b2, p->perm, p->items, p->item_ptr, p->partial, p->params, p->x, p->xc, p->blk_cur, p->blk_cand, p->scale,
                   p->state, p->ticket, p->fixed};
    for (void* d : dev) C.put(d, false);
    C.put(p->h_state, true);
    C.put(p->h_x, true);
    delete p;
}

// Work-item length: a multiple of 32 observations; short enough that every SM gets several warps of
// work on small problems, long enough that n_items stays <= ~32k on huge ones (the fold is serial
// in one CTA).
static int pick_item_len(int64_t n_obs, int sm_count)
{
    // ~2 work items per resident warp (3 CTAs x 8 warps per SM) whatever the problem size: small problems
    // still spread over the chip (32-observation items), huge ones keep the number of item partials
    // (and the fold over them) bounded.
    int64_t len = n_obs / ((int64_t)sm_count * 48);
    len = (len + 31) / 32 * 32;
    if (len < 32) len = 32;
    if (len > 8192) len = 8192;
    return (int)len;
}

static EvalArgs make_eval_args(sba_ba_problem* p, const double t[3], double d1, double d2, double huber, float* res, float* jac, double* blk_out)
{
    EvalArgs E;
    E.b1 = p->b1; E.b2 = p->b2; E.perm = p->perm; E.items = p->items; E.item_ptr = p->item_ptr;
    E.item_len = p->item_len; E.n_obs = (int)p->n_obs; E.n_obs_dev = p->n_obs_dev;
    E.n_items = p->n_items; E.n_cam = p->n_cam; E.params = p->params; E.partial = p->partial;
    E.blk_out = blk_out; E.ticket = p->ticket; E.res = res; E.jac = jac; E.done = nullptr; E.tvec = nullptr; E.tran = 0; E.comm = p->comm_dev;
    E.k.t[0] = t[0]; E.k.t[1] = t[1]; E.k.t[2] = t[2];
    E.k.d1 = d1; E.k.d2 = d2; E.k.huber = huber;
    return E;
}

static LMArrays make_lm_arrays(sba_ba_problem* p)
{
    LMArrays A;
    A.x = p->x; A.xc = p->xc; A.blk_cur = p->blk_cur; A.blk_cand = p->blk_cand; A.scale = p->scale;
    A.params = p->params; A.st = p->state; A.n_cam = p->n_cam; A.d1 = 1.0; A.tran = 0;
    return A;
}

static int upload_rotations(sba_ba_problem* p, const double* r, double* dst)
{
    memcpy(p->h_x, r, (size_t)p->n_cam * 3 * sizeof(double));
    SBA_CUDA(cudaMemcpyAsync(dst, p->h_x, (size_t)p->n_cam * 3 * sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
    return SBA_OK;
}

// One evaluation: the kernel, then (many cameras) the stand-alone fold.  `decide` asks for the fused
// in-kernel LM decision when the problem allows it; returns whether it was fused.
constexpr int FOLD_IN_KERNEL_MAX_CAMS = 64;

template <bool WRITE>
static int launch_eval(sba_ba_problem* p, const EvalArgs& E, const LMArrays& A, bool decide, bool* fused)
{
    cudaStream_t st = p->ctx->stream;
    const bool big = p->n_cam > FOLD_IN_KERNEL_MAX_CAMS;
    const bool fuse = decide && !big && !p->allreduce;   // a peer-memory comm does NOT prevent fusing: the exchange runs inside the kernel
    prof_begin(p->ctx, SBA_KERNEL_BA_EVAL);
    const dim3 grid(p->eval_blocks), block(EVAL_THREADS);
    if (E.n_obs_dev && !WRITE && !E.tran && !big) {
        if (fuse) ba_rot_eval_kernel<false, 2, false, true><<<grid, block, 0, st>>>(E, A);
        else ba_rot_eval_kernel<false, 1, false, true><<<grid, block, 0, st>>>(E, A);
    } else if (E.n_obs_dev) {
        sba::set_error("device-side observation count is only supported for fused single-camera rotation solves");
        return SBA_ERR_UNSUPPORTED;
    } else if (E.tran) {
        if (big) ba_rot_eval_kernel<WRITE, 0, true, false><<<grid, block, 0, st>>>(E, A);
        else if (fuse) ba_rot_eval_kernel<WRITE, 2, true, false><<<grid, block, 0, st>>>(E, A);
        else ba_rot_eval_kernel<WRITE, 1, true, false><<<grid, block, 0, st>>>(E, A);
    } else {
        if (big) ba_rot_eval_kernel<WRITE, 0, false, false><<<grid, block, 0, st>>>(E, A);
        else if (fuse) ba_rot_eval_kernel<WRITE, 2, false, false><<<grid, block, 0, st>>>(E, A);
        else ba_rot_eval_kernel<WRITE, 1, false, false><<<grid, block, 0, st>>>(E, A);
    }
    prof_end(p->ctx, SBA_KERNEL_BA_EVAL);
    SBA_LAUNCHED(p->ctx);
    if (big) {
        ba_fold_kernel<<<(p->n_cam + EVAL_WARPS - 1) / EVAL_WARPS, EVAL_THREADS, 0, st>>>(E.partial, E.item_ptr, p->n_cam, E.params, E.k.d1, E.tran, E.blk_out, E.done);
        SBA_LAUNCHED(p->ctx);
        if (p->comm_dev) {
            const int count = p->n_cam * 10;
            const int slice_len = std::max(512, (count + sba::COMM_MAX_SLICES - 1) / sba::COMM_MAX_SLICES);
            ba_exchange_kernel<<<(count + slice_len - 1) / slice_len, EVAL_THREADS, 0, st>>>(*p->comm_host, E.blk_out, count, slice_len, E.done);
            SBA_LAUNCHED(p->ctx);
        }
    }
    SBA_CUDA(cudaGetLastError());
    if (fused) *fused = fuse;
    return SBA_OK;
}

extern "C" {

}  // extern "C"

namespace sba {
int ba_problem_create_impl(sba_ctx* c, const float* b1, const float* b2, const int32_t* cam, int64_t n_obs, int n_cam, int mem, bool borrow,
                           const int* d_n_obs, sba_ba_problem** out);
}

extern "C" {

int sba_ba_problem_create(sba_ctx* c, const float* b1, const float* b2, const int32_t* cam, int64_t n_obs, int n_cam, int mem,
                          sba_ba_problem** out)
{
    return sba::ba_problem_create_impl(c, b1, b2, cam, n_obs, n_cam, mem, false, nullptr, out);
}

}  // extern "C"

// `borrow` (device pointers, no camera sort): the problem reads the caller's b1/b2 in place; they must
// stay valid and unchanged until the problem is destroyed.
// `d_n_obs` (borrowed single-camera problems only): the real observation count is read from device
// memory by the kernels and n_obs is just the capacity -- lets a caller enqueue the whole solve without
// first fetching the count to the host.
int sba::ba_problem_create_impl(sba_ctx* c, const float* b1, const float* b2, const int32_t* cam, int64_t n_obs, int n_cam, int mem,
                                bool borrow, const int* d_n_obs, sba_ba_problem** out)
{
    SBA_CHECK_ARG(c && out && n_obs >= 0 && n_cam >= 1 && n_obs < ((int64_t)1 << 31));
    SBA_CHECK_ARG(n_obs == 0 || (b1 && b2));
    *out = nullptr;
    SBA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    sba_ba_problem* p = new sba_ba_problem();
    p->ctx = c; p->n_obs = n_obs; p->n_cam = n_cam;
    int status = SBA_OK;
    auto fail = [&](int s) { free_problem(p); return s; };
#define P_CUDA(call)                                                                              \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            sba::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return fail(SBA_ERR_CUDA);                                                            \
        }                                                                                         \
    } while (0)

    size_t nb = (size_t)(n_obs ? n_obs : 1) * sizeof(float4);
    const bool need_sort = (cam != nullptr && n_cam > 1 && n_obs > 0);
    p->borrowed = borrow && !need_sort && mem == SBA_MEM_DEVICE && n_obs > 0;
    if (p->borrowed && n_cam == 1) p->n_obs_dev = d_n_obs;
    if (p->borrowed) {
        p->b1 = (float4*)b1;
        p->b2 = (float4*)b2;
    } else {
        P_CUDA(c->cache.get((void**)&p->b1, nb, false));
        P_CUDA(c->cache.get((void**)&p->b2, nb, false));
    }
    cudaMemcpyKind kind = mem == SBA_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    std::vector<int> counts(n_cam, 0);
    if (!need_sort) {
        if (n_obs && !p->borrowed) {
            P_CUDA(cudaMemcpyAsync(p->b1, b1, nb, kind, st));
            P_CUDA(cudaMemcpyAsync(p->b2, b2, nb, kind, st));
        }
        counts[0] = (int)n_obs;
        if (cam != nullptr && n_cam == 1) { /* ids must all be 0; trusted */ }
    } else {
        // group observations by camera: stable radix sort of (cam, index), then gather
        const float *d_b1, *d_b2;
        const int32_t* d_cam;
        if ((status = stage_in(c, b1, (size_t)4 * n_obs, mem, SCR_IN0, &d_b1)) != SBA_OK) return fail(status);
        if ((status = stage_in(c, b2, (size_t)4 * n_obs, mem, SCR_IN1, &d_b2)) != SBA_OK) return fail(status);
        if ((status = stage_in(c, cam, (size_t)n_obs, mem, SCR_IN2, &d_cam)) != SBA_OK) return fail(status);
        if ((status = c->scratch[SCR_WORK0].ensure((size_t)n_obs * 4, st)) != SBA_OK) return fail(status);
        if ((status = c->scratch[SCR_WORK1].ensure((size_t)n_obs * 4, st)) != SBA_OK) return fail(status);
        if ((status = c->scratch[SCR_WORK3].ensure((size_t)(n_cam + 1) * 4, st)) != SBA_OK) return fail(status);
        int32_t* d_idx = c->scratch[SCR_WORK0].as<int32_t>();
        int32_t* d_keys_sorted = c->scratch[SCR_WORK1].as<int32_t>();
        int* d_counts = c->scratch[SCR_WORK3].as<int>();
        P_CUDA(c->cache.get((void**)&p->perm, (size_t)n_obs * 4, false));
        int gb = (int)std::min<int64_t>(ceil_div64(n_obs, 256), (int64_t)c->sm_count * 8);
        iota_kernel<<<gb, 256, 0, st>>>(d_idx, n_obs);
        SBA_LAUNCHED(c);
        P_CUDA(cudaMemsetAsync(d_counts, 0, (size_t)(n_cam + 1) * 4, st));
        cam_hist_kernel<<<gb, 256, 0, st>>>(d_cam, n_obs, n_cam, d_counts, d_counts + n_cam);
        SBA_LAUNCHED(c);
        int end_bit = 1;
        while ((1ll << end_bit) < n_cam) end_bit++;
        size_t tmp_bytes = 0;
        P_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_cam, d_keys_sorted, d_idx, p->perm, (int)n_obs, 0, end_bit, st));
        if ((status = c->scratch[SCR_WORK2].ensure(tmp_bytes, st)) != SBA_OK) return fail(status);
        P_CUDA(cub::DeviceRadixSort::SortPairs(c->scratch[SCR_WORK2].p, tmp_bytes, d_cam, d_keys_sorted, d_idx, p->perm, (int)n_obs, 0, end_bit, st));
        SBA_LAUNCHED(c);
        gather_obs_kernel<<<gb, 256, 0, st>>>((const float4*)d_b1, (const float4*)d_b2, p->perm, n_obs, p->b1, p->b2);
        SBA_LAUNCHED(c);
        std::vector<int> hc(n_cam + 1);
        P_CUDA(cudaMemcpyAsync(hc.data(), d_counts, (size_t)(n_cam + 1) * 4, cudaMemcpyDeviceToHost, st));
        P_CUDA(cudaStreamSynchronize(st));
        if (hc[n_cam]) {
            sba::set_error("camera index out of range [0, %d)", n_cam);
            return fail(SBA_ERR_INVALID);
        }
        for (int k = 0; k < n_cam; k++) counts[k] = hc[k];
    }

    // work items: arithmetic (uniform) for a single camera, a host-built table otherwise
    const int len = pick_item_len(n_obs, c->sm_count);
    p->item_len = len;
    const bool uniform = (n_cam == 1);
    std::vector<Item> items;
    std::vector<int> item_ptr(n_cam + 1, 0);
    int64_t off = 0;
    for (int k = 0; k < n_cam && !uniform; k++) {
        item_ptr[k] = (int)items.size();
        for (int64_t s = 0; s < counts[k]; s += len) {
            Item it;
            it.start = off + s;
            it.count = (int)std::min<int64_t>(len, counts[k] - s);
            it.cam = k;
            items.push_back(it);
        }
        off += counts[k];
    }
    item_ptr[n_cam] = (int)items.size();
    p->n_items = uniform ? (int)((n_obs + len - 1) / len) : (int)items.size();
    size_t ni = p->n_items ? p->n_items : 1;
    if (!uniform) {
        P_CUDA(c->cache.get((void**)&p->items, ni * sizeof(Item), false));
        P_CUDA(c->cache.get((void**)&p->item_ptr, (size_t)(n_cam + 1) * sizeof(int), false));
    }
    P_CUDA(c->cache.get((void**)&p->params, (size_t)n_cam * sizeof(CamParams), false));
    P_CUDA(c->cache.get((void**)&p->x, (size_t)n_cam * 3 * sizeof(double), false));
    P_CUDA(c->cache.get((void**)&p->xc, (size_t)n_cam * 3 * sizeof(double), false));
    P_CUDA(c->cache.get((void**)&p->blk_cur, (size_t)n_cam * 10 * sizeof(double), false));
    P_CUDA(c->cache.get((void**)&p->blk_cand, (size_t)n_cam * 10 * sizeof(double), false));
    P_CUDA(c->cache.get((void**)&p->scale, (size_t)n_cam * 3 * sizeof(double), false));
    P_CUDA(c->cache.get((void**)&p->state, sizeof(LMState), false));
    P_CUDA(c->cache.get((void**)&p->ticket, sizeof(unsigned int), false));
    P_CUDA(c->cache.get((void**)&p->h_state, sizeof(LMState), true));
    P_CUDA(c->cache.get((void**)&p->h_x, (size_t)n_cam * 3 * sizeof(double), true));
    P_CUDA(cudaMemsetAsync(p->ticket, 0, sizeof(unsigned int), st));
    if (!uniform) {
        if (!items.empty()) P_CUDA(cudaMemcpyAsync(p->items, items.data(), items.size() * sizeof(Item), cudaMemcpyHostToDevice, st));
        P_CUDA(cudaMemcpyAsync(p->item_ptr, item_ptr.data(), (size_t)(n_cam + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    }
    // host vectors / the caller's host buffers are released after this; nothing to wait for when
    // everything already lives on the device
    if (!uniform || mem == SBA_MEM_HOST) P_CUDA(cudaStreamSynchronize(st));

    // grid: enough CTAs for every work item's warp, capped at 2 resident CTAs per SM x 4 waves
    int want = (p->n_items + EVAL_WARPS - 1) / EVAL_WARPS;
    if (want < 1) want = 1;
    p->eval_blocks = std::min(want, c->sm_count * 8);
    // uniform layout: one partial per CTA; otherwise one per work item
    P_CUDA(c->cache.get((void**)&p->partial, (size_t)std::max<size_t>(ni, (size_t)p->eval_blocks) * NMOM * sizeof(double), false));
    *out = p;
    return SBA_OK;
#undef P_CUDA
}

extern "C" {

int sba_ba_problem_destroy(sba_ba_problem* p)
{
    if (!p) return SBA_OK;
    // Blocks go back to the context's cache and are only ever reused by work enqueued later on the
    // same stream, so no synchronisation is needed here.
    free_problem(p);
    return SBA_OK;
}

int sba_ba_problem_set_allreduce(sba_ba_problem* p, sba_allreduce_fn fn, void* user)
{
    SBA_CHECK_ARG(p != nullptr);
    p->allreduce = fn;
    p->allreduce_user = user;
    return SBA_OK;
}

}  // extern "C"

// Host side of the peer-memory exchange.  Each rank's buffer: [64 per-slice sequence flags][2 slots x stride doubles].
struct sba_comm {
    sba_ctx* ctx = nullptr;
    int rank = 0, world = 1, stride = 0;
    uint8_t* local = nullptr;
    void* peer_base[sba::COMM_MAX_WORLD] = {};
    sba::CommDev host{};
    sba::CommDev* dev = nullptr;
    unsigned long long* seq = nullptr;
    bool connected = false;
};

extern "C" {

int sba_comm_create(sba_ctx* c, int rank, int world, int max_cameras, sba_comm** out, void* ipc_handle_out)
{
    SBA_CHECK_ARG(c && out && ipc_handle_out && world >= 1 && world <= sba::COMM_MAX_WORLD && rank >= 0 && rank < world && max_cameras >= 1);
    static_assert(sizeof(cudaIpcMemHandle_t) == SBA_COMM_HANDLE_BYTES, "IPC handle size");
    SBA_CUDA(cudaSetDevice(c->device));
    sba_comm* m = new sba_comm();
    m->ctx = c; m->rank = rank; m->world = world;
    m->stride = (max_cameras * 10 + 15) / 16 * 16;
    const size_t bytes = sba::COMM_HEADER_BYTES + (size_t)2 * m->stride * sizeof(double);
    cudaIpcMemHandle_t h;
    if (cudaMalloc((void**)&m->local, bytes) != cudaSuccess || cudaMemset(m->local, 0, bytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&h, m->local) != cudaSuccess) {
        sba::set_error("sba_comm_create: %s", cudaGetErrorString(cudaGetLastError()));
        if (m->local) cudaFree(m->local);
        delete m;
        return SBA_ERR_COMM;
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *out = m;
    return SBA_OK;
}

int sba_comm_connect(sba_comm* m, const void* ipc_handles)
{
    SBA_CHECK_ARG(m && ipc_handles && !m->connected);
    SBA_CUDA(cudaSetDevice(m->ctx->device));
    for (int r = 0; r < m->world; r++) {
        if (r == m->rank) m->peer_base[r] = m->local;
        else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const uint8_t*)ipc_handles + (size_t)r * SBA_COMM_HANDLE_BYTES, sizeof(h));
            cudaError_t e = cudaIpcOpenMemHandle(&m->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                sba::set_error("cudaIpcOpenMemHandle(rank %d): %s (peers must be GPUs of one NVLink/PCIe-P2P box)", r, cudaGetErrorString(e));
                return SBA_ERR_COMM;
            }
        }
        m->host.flag[r] = (volatile unsigned long long*)m->peer_base[r];
        m->host.data[r] = (double*)((uint8_t*)m->peer_base[r] + sba::COMM_HEADER_BYTES);
    }
    m->host.rank = m->rank; m->host.world = m->world; m->host.stride = m->stride;
    SBA_CUDA(cudaMalloc((void**)&m->seq, 2 * sizeof(unsigned long long)));
    SBA_CUDA(cudaMemset(m->seq, 0, 2 * sizeof(unsigned long long)));
    m->host.seq = m->seq;
    m->host.ticket = (unsigned int*)(m->seq + 1);
    SBA_CUDA(cudaMalloc((void**)&m->dev, sizeof(sba::CommDev)));
    SBA_CUDA(cudaMemcpy(m->dev, &m->host, sizeof(sba::CommDev), cudaMemcpyHostToDevice));
    m->connected = true;
    return SBA_OK;
}

int sba_comm_destroy(sba_comm* m)
{
    if (!m) return SBA_OK;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    for (int r = 0; r < m->world; r++)
        if (r != m->rank && m->peer_base[r]) cudaIpcCloseMemHandle(m->peer_base[r]);
    if (m->dev) cudaFree(m->dev);
    if (m->seq) cudaFree(m->seq);
    if (m->local) cudaFree(m->local);
    delete m;
    return SBA_OK;
}

int sba_ba_problem_set_comm(sba_ba_problem* p, sba_comm* m)
{
    SBA_CHECK_ARG(p != nullptr);
    if (!m) { p->comm_dev = nullptr; p->comm_host = nullptr; return SBA_OK; }
    SBA_CHECK_ARG(m->connected && p->n_cam * 10 <= m->stride && m->ctx == p->ctx);
    p->comm_dev = m->dev;
    p->comm_host = &m->host;
    p->allreduce = nullptr;
    return SBA_OK;
}

int sba_ba_rot_eval(sba_ba_problem* p, const double* r, const double t[3], double d1, double d2, double huber, float* res, float* jac,
                    double* H, double* g, double* cost, int mem)
{
    SBA_CHECK_ARG(p && r && t);
    sba_ctx* c = p->ctx;
    SBA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int n_cam = p->n_cam;
    SBA_CUDA(cudaStreamSynchronize(st));  // h_x reuse
    SBA_TRY(upload_rotations(p, r, p->xc));
    ba_cam_params_kernel<<<(n_cam + 127) / 128, 128, 0, st>>>(p->xc, n_cam, d1, p->params);
    SBA_LAUNCHED(c);
    float *d_res, *d_jac;
    double *d_H, *d_g, *d_cost;
    SBA_TRY(stage_out(c, res, (size_t)3 * p->n_obs, mem, SCR_OUT0, &d_res));
    SBA_TRY(stage_out(c, jac, (size_t)9 * p->n_obs, mem, SCR_OUT1, &d_jac));
    SBA_TRY(stage_out(c, H, (size_t)6 * n_cam, mem, SCR_OUT2, &d_H));
    SBA_TRY(stage_out(c, g, (size_t)3 * n_cam, mem, SCR_OUT3, &d_g));
    SBA_TRY(stage_out(c, cost, (size_t)n_cam, mem, SCR_OUT4, &d_cost));
    EvalArgs E = make_eval_args(p, t, d1, d2, huber, d_res, d_jac, p->blk_cand);
    LMArrays A = make_lm_arrays(p);
    A.d1 = d1;
    if (d_res || d_jac) SBA_TRY(launch_eval<true>(p, E, A, false, nullptr));
    else SBA_TRY(launch_eval<false>(p, E, A, false, nullptr));
    if (p->allreduce) {
        if (p->allreduce(p->blk_cand, (int64_t)n_cam * 10, p->allreduce_user) != 0) {
            sba::set_error("allreduce callback failed");
            return SBA_ERR_COMM;
        }
    }
    ba_unpack_blocks_kernel<<<(n_cam + 127) / 128, 128, 0, st>>>(p->blk_cand, n_cam, d_H, d_g, d_cost);
    SBA_LAUNCHED(c);
    SBA_TRY(copy_out(c, res, d_res, (size_t)3 * p->n_obs, mem));
    SBA_TRY(copy_out(c, jac, d_jac, (size_t)9 * p->n_obs, mem));
    SBA_TRY(copy_out(c, H, d_H, (size_t)6 * n_cam, mem));
    SBA_TRY(copy_out(c, g, d_g, (size_t)3 * n_cam, mem));
    SBA_TRY(copy_out(c, cost, d_cost, (size_t)n_cam, mem));
    return finish(c, mem);
}

}  // extern "C"

namespace sba {

BaView ba_problem_view(sba_ba_problem* p)
{
    return BaView{p->ctx, p->b1, p->b2, p->n_obs, p->n_cam, p->n_obs_dev};
}

constexpr int LM_CHUNK = 6;   // evaluations enqueued between two looks at the solver state

// (1) host side only: starting point and a fresh solver state into the problem's pinned mailboxes.
void ba_solve_prepare_host(sba_ba_problem* p, const double* r0, int max_iter)
{
    memcpy(p->h_x, r0, (size_t)p->n_cam * 3 * sizeof(double));
    LMState init{};
    init.radius = 1e4; init.dec_factor = 2.0; init.max_iter = max_iter; init.phase = 0;
    *p->h_state = init;
}

static int enqueue_chunk(sba_ba_problem* p, const EvalArgs& E, const LMArrays& A, int n)
{
    sba_ctx* c = p->ctx;
    cudaStream_t st = c->stream;
    for (int k = 0; k < n; k++) {
        bool fused = false;
        SBA_TRY(launch_eval<false>(p, E, A, true, &fused));
        if (!fused) {
            if (p->allreduce && p->allreduce(p->blk_cand, (int64_t)p->n_cam * 10, p->allreduce_user) != 0) {
                sba::set_error("allreduce callback failed");
                return SBA_ERR_COMM;
            }
            ba_decide_kernel<<<1, EVAL_THREADS, 0, st>>>(A);
            SBA_LAUNCHED(c);
        }
    }
    SBA_CUDA(cudaMemcpyAsync(p->h_state, p->state, sizeof(LMState), cudaMemcpyDeviceToHost, st));
    // small problems: fetch the parameters with the state (one round trip per chunk)
    if (p->n_cam <= 64) SBA_CUDA(cudaMemcpyAsync(p->h_x, p->x, (size_t)p->n_cam * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    return SBA_OK;
}

// (2) stream side, no synchronisation: mailboxes -> device, rotation
// tables, the first chunk of evaluations, state (+ parameters) back to the mailboxes.
// Evaluations that start after convergence return immediately (state.done).
// Evaluation / LM arguments of a solve in either mode (rotation free: t uniform and fixed;
// translation free: the rotations in p->fixed are fixed and t comes from the candidate vector).
static void solve_args(sba_ba_problem* p, const double t[3], double d1, double d2, double huber, bool tran, EvalArgs* E, LMArrays* A)
{
    const double zero[3] = {0, 0, 0};
    *E = make_eval_args(p, tran ? zero : t, d1, d2, huber, nullptr, nullptr, p->blk_cand);
    E->done = &p->state->done;
    *A = make_lm_arrays(p);
    A->d1 = d1;
    if (tran) { E->tvec = p->xc; E->tran = 1; A->tran = 1; }
}

int ba_solve_enqueue(sba_ba_problem* p, const double t[3], double d1, double d2, double huber, int max_iter, int* launched, bool tran)
{
    sba_ctx* c = p->ctx;
    cudaStream_t st = c->stream;
    const int n_cam = p->n_cam;
    SBA_CUDA(cudaMemcpyAsync(p->x, p->h_x, (size_t)n_cam * 3 * sizeof(double), cudaMemcpyHostToDevice, st));
    SBA_CUDA(cudaMemcpyAsync(p->xc, p->x, (size_t)n_cam * 3 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    SBA_CUDA(cudaMemcpyAsync(p->state, p->h_state, sizeof(LMState), cudaMemcpyHostToDevice, st));
    // rotation tables: from the starting rotations, or once from the fixed ones
    ba_cam_params_kernel<<<(n_cam + 127) / 128, 128, 0, st>>>(tran ? p->fixed : p->xc, n_cam, d1, p->params);
    SBA_LAUNCHED(c);
    EvalArgs E;
    LMArrays A;
    solve_args(p, t, d1, d2, huber, tran, &E, &A);
    const int n = std::min(LM_CHUNK, max_iter + 1);
    SBA_TRY(enqueue_chunk(p, E, A, n));
    *launched = n;
    return SBA_OK;
}

// Fused alternative to (2) for the pair pipeline: keypoints + match list in, the whole solve enqueued as ONE launch.
int ba_pair_solve_enqueue(sba_ba_problem* p, const float* key_l, const float* key_r, const int32_t* qi, const int32_t* ti, const int32_t* d_n,
                          int cap, int cs, int w, int h, const double t[3], double d1, double d2, double huber, int max_iter, int* launched)
{
    sba_ctx* c = p->ctx;
    cudaStream_t st = c->stream;
    PairSolveArgs P;
    for (int k = 0; k < 3; k++) P.r0[k] = p->h_x[k];    // ba_solve_prepare_host put the starting point and the fresh state there
    P.init = *p->h_state;
    P.h_state = p->h_state; P.h_x = p->h_x;
    p->h_state->done = 0;
    P.key_l = (const float2*)key_l; P.key_r = (const float2*)key_r; P.qi = qi; P.ti = ti; P.d_n = d_n;
    P.cap = cap; P.cs = cs; P.w = w; P.h = h; P.b1 = p->b1; P.b2 = p->b2;
    P.k.t[0] = t[0]; P.k.t[1] = t[1]; P.k.t[2] = t[2]; P.k.d1 = d1; P.k.d2 = d2; P.k.huber = huber;
    LMArrays A = make_lm_arrays(p);
    A.d1 = d1;
    // one cluster: 16 CTAs (needs the non-portable cluster size attribute), else the portable 8
    static int cluster_size = 0;
    if (cluster_size == 0) {
        cluster_size = 8;
        if (cudaFuncSetAttribute(ba_pair_solve_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
            cudaLaunchConfig_t probe = {};
            cudaLaunchAttribute pa[1];
            pa[0].id = cudaLaunchAttributeClusterDimension;
            pa[0].val.clusterDim.x = PAIR_SOLVE_CLUSTER; pa[0].val.clusterDim.y = 1; pa[0].val.clusterDim.z = 1;
            probe.gridDim = dim3(PAIR_SOLVE_CLUSTER); probe.blockDim = dim3(PAIR_SOLVE_THREADS); probe.attrs = pa; probe.numAttrs = 1;
            int n_clusters = 0;
            if (cudaOccupancyMaxActiveClusters(&n_clusters, ba_pair_solve_kernel, &probe) == cudaSuccess && n_clusters >= 1) cluster_size = PAIR_SOLVE_CLUSTER;
        }
        cudaGetLastError();
    }
    const int grid = std::max(1, std::min(cluster_size, (cap + PAIR_SOLVE_THREADS - 1) / PAIR_SOLVE_THREADS));   // small pairs: a smaller cluster
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)grid; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(PAIR_SOLVE_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = attr; cfg.numAttrs = (c->pdl || c->pdl_small) ? 2 : 1;
    prof_begin(c, SBA_KERNEL_BA_EVAL);
    SBA_CUDA(cudaLaunchKernelEx(&cfg, ba_pair_solve_kernel, P, A));
    prof_end(c, SBA_KERNEL_BA_EVAL);
    SBA_LAUNCHED(c);
    SBA_CUDA(cudaGetLastError());
    *launched = max_iter + 1;   // the kernel runs the solve to completion: ba_solve_finish has nothing to add
    return SBA_OK;
}

// (3) wait for the enqueued chunk; keep going chunk by chunk until the solver reports done.
int ba_solve_finish(sba_ba_problem* p, double* r_out, const double t[3], double d1, double d2, double huber, int max_iter, int launched,
                    sba_solve_summary* summary, bool tran)
{
    sba_ctx* c = p->ctx;
    cudaStream_t st = c->stream;
    const int n_cam = p->n_cam;
    const int max_evals = max_iter + 1;
    SBA_CUDA(cudaStreamSynchronize(st));
    if (!p->h_state->done && launched < max_evals) {
        EvalArgs E;
        LMArrays A;
        solve_args(p, t, d1, d2, huber, tran, &E, &A);
        while (!p->h_state->done && launched < max_evals) {
            const int n = std::min(LM_CHUNK, max_evals - launched);
            SBA_TRY(enqueue_chunk(p, E, A, n));
            launched += n;
            SBA_CUDA(cudaStreamSynchronize(st));
        }
    }
    if (n_cam > 64) {
        SBA_CUDA(cudaMemcpyAsync(p->h_x, p->x, (size_t)n_cam * 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
        SBA_CUDA(cudaStreamSynchronize(st));
    }
    memcpy(r_out, p->h_x, (size_t)n_cam * 3 * sizeof(double));
    if (summary) {
        const LMState& S = *p->h_state;
        summary->iterations = S.iter;
        summary->num_successful = S.num_successful;
        summary->termination = S.done ? S.termination : 0;
        summary->evaluations = S.evals;
        summary->initial_cost = S.initial_cost;
        summary->final_cost = S.cost;
        summary->final_radius = S.radius;
    }
    return SBA_OK;
}

}  // namespace sba

extern "C" {

int sba_ba_rot_solve(sba_ba_problem* p, double* r_inout, const double t[3], double d1, double d2, double huber, int max_iter,
                     sba_solve_summary* summary)
{
    SBA_CHECK_ARG(p && r_inout && t && max_iter >= 0);
    SBA_CUDA(cudaSetDevice(p->ctx->device));
    // h_x / h_state are pinned mailboxes of this problem; every earlier use ended with a synchronise
    ba_solve_prepare_host(p, r_inout, max_iter);
    int launched = 0;
    SBA_TRY(ba_solve_enqueue(p, t, d1, d2, huber, max_iter, &launched, false));
    return ba_solve_finish(p, r_inout, t, d1, d2, huber, max_iter, launched, summary, false);
}

static int upload_fixed(sba_ba_problem* p, const double* r_fixed)
{
    sba_ctx* c = p->ctx;
    const size_t bytes = (size_t)p->n_cam * 3 * sizeof(double);
    if (!p->fixed) SBA_CUDA(c->cache.get((void**)&p->fixed, bytes, false));
    SBA_CUDA(cudaStreamSynchronize(c->stream));   // the pinned mailbox h_x is reused as the staging buffer
    memcpy(p->h_x, r_fixed, bytes);
    SBA_CUDA(cudaMemcpyAsync(p->fixed, p->h_x, bytes, cudaMemcpyHostToDevice, c->stream));
    SBA_CUDA(cudaStreamSynchronize(c->stream));
    return SBA_OK;
}

int sba_ba_tran_solve(sba_ba_problem* p, const double* r_fixed, double* t_inout, double d1, double d2, double huber, int max_iter,
                      sba_solve_summary* summary)
{
    SBA_CHECK_ARG(p && r_fixed && t_inout && max_iter >= 0);
    SBA_CUDA(cudaSetDevice(p->ctx->device));
    SBA_TRY(upload_fixed(p, r_fixed));
    ba_solve_prepare_host(p, t_inout, max_iter);
    const double unused[3] = {0, 0, 0};
    int launched = 0;
    SBA_TRY(ba_solve_enqueue(p, unused, d1, d2, huber, max_iter, &launched, true));
    return ba_solve_finish(p, t_inout, unused, d1, d2, huber, max_iter, launched, summary, true);
}

int sba_ba_tran_eval(sba_ba_problem* p, const double* r_fixed, const double* tv, double d1, double d2, double huber, float* res, double* H,
                     double* g, double* cost, int mem)
{
    SBA_CHECK_ARG(p && r_fixed && tv);
    sba_ctx* c = p->ctx;
    SBA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int n_cam = p->n_cam;
    SBA_TRY(upload_fixed(p, r_fixed));
    SBA_TRY(upload_rotations(p, tv, p->xc));   // the translations ride in the candidate vector
    ba_cam_params_kernel<<<(n_cam + 127) / 128, 128, 0, st>>>(p->fixed, n_cam, d1, p->params);
    SBA_LAUNCHED(c);
    float* d_res;
    double *d_H, *d_g, *d_cost;
    SBA_TRY(stage_out(c, res, (size_t)3 * p->n_obs, mem, SCR_OUT0, &d_res));
    SBA_TRY(stage_out(c, H, (size_t)6 * n_cam, mem, SCR_OUT2, &d_H));
    SBA_TRY(stage_out(c, g, (size_t)3 * n_cam, mem, SCR_OUT3, &d_g));
    SBA_TRY(stage_out(c, cost, (size_t)n_cam, mem, SCR_OUT4, &d_cost));
    const double zero[3] = {0, 0, 0};
    EvalArgs E = make_eval_args(p, zero, d1, d2, huber, d_res, nullptr, p->blk_cand);
    E.tvec = p->xc;
    E.tran = 1;
    LMArrays A = make_lm_arrays(p);
    A.d1 = d1;
    A.tran = 1;
    if (d_res) SBA_TRY(launch_eval<true>(p, E, A, false, nullptr));
    else SBA_TRY(launch_eval<false>(p, E, A, false, nullptr));
    if (p->allreduce && p->allreduce(p->blk_cand, (int64_t)n_cam * 10, p->allreduce_user) != 0) {
        sba::set_error("allreduce callback failed");
        return SBA_ERR_COMM;
    }
    ba_unpack_blocks_kernel<<<(n_cam + 127) / 128, 128, 0, st>>>(p->blk_cand, n_cam, d_H, d_g, d_cost);
    SBA_LAUNCHED(c);
    SBA_TRY(copy_out(c, res, d_res, (size_t)3 * p->n_obs, mem));
    SBA_TRY(copy_out(c, H, d_H, (size_t)6 * n_cam, mem));
    SBA_TRY(copy_out(c, g, d_g, (size_t)3 * n_cam, mem));
    SBA_TRY(copy_out(c, cost, d_cost, (size_t)n_cam, mem));
    return finish(c, mem);
}

int sba_ba_rot_eval_timed(sba_ba_problem* p, const double* r, const double t[3], double d1, double d2, double huber, int materialise,
                          int iters, float* mean_ms)
{
    SBA_CHECK_ARG(p && r && t && iters > 0 && mean_ms);
    sba_ctx* c = p->ctx;
    SBA_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int n_cam = p->n_cam;
    SBA_CUDA(cudaStreamSynchronize(st));
    SBA_TRY(upload_rotations(p, r, p->xc));
    ba_cam_params_kernel<<<(n_cam + 127) / 128, 128, 0, st>>>(p->xc, n_cam, d1, p->params);
    SBA_LAUNCHED(c);
    float *d_res = nullptr, *d_jac = nullptr;
    if (materialise) {
        SBA_TRY(c->scratch[SCR_OUT0].ensure((size_t)3 * p->n_obs * sizeof(float) + 16, st));
        SBA_TRY(c->scratch[SCR_OUT1].ensure((size_t)9 * p->n_obs * sizeof(float) + 16, st));
        d_res = c->scratch[SCR_OUT0].as<float>();
        d_jac = c->scratch[SCR_OUT1].as<float>();
    }
    EvalArgs E = make_eval_args(p, t, d1, d2, huber, d_res, d_jac, p->blk_cand);
    LMArrays A = make_lm_arrays(p);
    A.d1 = d1;
    cudaEvent_t e0, e1;
    SBA_CUDA(cudaEventCreate(&e0));
    SBA_CUDA(cudaEventCreate(&e1));
    for (int k = 0; k < 3; k++) {
        if (materialise) SBA_TRY(launch_eval<true>(p, E, A, false, nullptr));
        else SBA_TRY(launch_eval<false>(p, E, A, false, nullptr));
    }
    SBA_CUDA(cudaEventRecord(e0, st));
    for (int k = 0; k < iters; k++) {
        if (materialise) SBA_TRY(launch_eval<true>(p, E, A, false, nullptr));
        else SBA_TRY(launch_eval<false>(p, E, A, false, nullptr));
    }
    SBA_CUDA(cudaEventRecord(e1, st));
    SBA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    SBA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *mean_ms = ms / iters;
    return SBA_OK;
}

}  // extern "C"

#ifdef SBA_TC_TRACE
extern "C" int sba_ps_trace_read(unsigned long long* out /* [64] */)
{
    return cudaMemcpyFromSymbol(out, sba::g_ps_trace, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : -1;
}
#endif
