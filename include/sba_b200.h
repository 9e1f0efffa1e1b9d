/*
 * sba_b200.h -- C ABI of libsba_b200.so: the B200 (sm_100a) implementation of the hot path of
 * whdlgp/spherical_bundle_adjuster.
 *
 * The reference has no FFI layer; its boundary is the public C++ class API
 * (equi2cube.hpp:20-32, equi2cube_surf.hpp:7-18, feature_matcher.hpp:24-49,
 * spherical_bundle_adjuster.hpp:15-23,86-115).  The drop-in facade classes in
 * spherical_bundle_adjuster_b200/host/ keep those signatures and call the entry points below;
 * a maintainer of the reference binds the same entry points directly.
 *
 * Conventions
 *   - every function returns 0 (SBA_OK) or a negative sba_status; sba_last_error() gives text.
 *   - plain pointers and sizes only.  `mem` says where the data pointers of THAT call live:
 *     SBA_MEM_HOST (the library stages through its own device buffers and copies results back,
 *     synchronising before it returns) or SBA_MEM_DEVICE (zero copy, asynchronous on the
 *     context's stream; scalars returned through host pointers force a synchronisation).
 *   - there is no CPU fallback: without a usable CUDA device sba_ctx_create fails.
 *   - a context is not thread safe; use one per host thread (the reference's classes are not
 *     re-entrant either: feature_matcher.hpp:44-48).
 */
#ifndef SBA_B200_H
#define SBA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBA_B200_VERSION 100 /* 0.1.0 */

typedef enum sba_status {
    SBA_OK = 0,
    SBA_ERR_INVALID = -1,     /* bad argument */
    SBA_ERR_CUDA = -2,        /* CUDA runtime/driver error (text in sba_last_error) */
    SBA_ERR_NOMEM = -3,
    SBA_ERR_UNSUPPORTED = -4, /* e.g. descriptor dimension not handled by the selected algorithm */
    SBA_ERR_NO_DEVICE = -5,
    SBA_ERR_COMM = -6
} sba_status;

typedef enum sba_mem { SBA_MEM_HOST = 0, SBA_MEM_DEVICE = 1 } sba_mem;

typedef struct sba_ctx sba_ctx;

int sba_version(void);
/* Text of the last error raised on this thread ("" if none). */
const char* sba_last_error(void);

/* Create a context on CUDA device `device`.  `stream` is a cudaStream_t (as void*) that all work
 * of this context is enqueued on, or NULL to let the context create its own non-blocking stream. */
int sba_ctx_create(int device, void* stream, sba_ctx** out);
int sba_ctx_destroy(sba_ctx* ctx);
int sba_ctx_set_stream(sba_ctx* ctx, void* stream);
void* sba_ctx_get_stream(sba_ctx* ctx);
int sba_ctx_synchronize(sba_ctx* ctx);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
int64_t sba_ctx_launch_count(sba_ctx* ctx);

/* Kernel timing for roofline reports.  When enabled, the dominant kernel of each stage is bracketed
 * by CUDA events on the context's stream; sba_ctx_kernel_ms returns the device time of the most
 * recent launch of that kernel (synchronising on its end event). */
typedef enum sba_kernel_id {
    SBA_KERNEL_MATCH = 0, /* the kNN distance kernel (SIMT or tensor-core) */
    SBA_KERNEL_REMAP = 1, /* the equi2cube gather */
    SBA_KERNEL_BA_EVAL = 2 /* the BA residual+Jacobian evaluation kernel */
} sba_kernel_id;
int sba_ctx_set_profiling(sba_ctx* ctx, int enable);
int sba_ctx_kernel_ms(sba_ctx* ctx, int kernel_id, float* ms);

/* ---------------------------------------------------------------------------------------------
 * equi2cube  (replaces equi2cube.cpp:12-302)
 *
 * erp      : n_images x h x w x 3 bytes, interleaved BGR, continuous (CV_8UC3).
 * strip_out: n_images x cube_size x (6*cube_size) x 3 bytes; faces in get_all's order
 *            left, front, right, back, top, bottom (equi2cube.cpp:293-298).
 * Source index per output pixel is the reference's truncating fp64 formula (equi2cube.cpp:26-48).
 * The one unchecked read of the reference (row == h at the exact bottom-centre pixel when theta
 * rounds to pi) is clamped to row h-1.
 * The index table depends only on (w, h, cube_size); it is built once per geometry (on the
 * device, with near-integer coordinates resolved by the host libm so every index is bit-exact)
 * and cached in the context.
 * ------------------------------------------------------------------------------------------- */
int sba_equi2cube(sba_ctx* ctx, const uint8_t* erp, int w, int h, int n_images, int cube_size, uint8_t* strip_out, int mem);

/* One face, cube_size x cube_size x 3 (get_left/get_front/get_right/get_back/get_top/get_bottom).
 * face: 0 left, 1 front, 2 right, 3 back, 4 top, 5 bottom. */
int sba_equi2cube_face(sba_ctx* ctx, const uint8_t* erp, int w, int h, int cube_size, int face, uint8_t* face_out, int mem);

/* Build (or fetch from the cache) the index table: lut_out[cube_size][6*cube_size] int32 source
 * pixel indices (row*w + col).  lut_out may be NULL to only warm the cache. */
/* The table-driven gathers (cube strips, spherical_surf bands) exist in two kernels: a direct gather and a tiled one
 * that stages each output tile's source bounding box in shared memory with bulk copies.  Which is faster depends on
 * the table and on the batch size, and is measured once when a plan is built (2 frames; a batch larger than L2).
 * A third kernel fetches the pixels of each 32 x 64 output tile in SOURCE order (a per-tile sorted form of the table), so
 * that one load instruction walks along a source row instead of across many.
 * mode 0 = use that choice, 1 = always direct, 2 = tiled / 3 = source-ordered wherever the geometry allows it (tests run all).
 * sba_remap_plan_info reports what the trials of a cube plan found: tiled_preferred bit 0 = small batches, bit 1 = large;
 * trial_ms = {small direct, small tiled, large direct, large tiled}. */
int sba_ctx_set_remap_kernel(sba_ctx* ctx, int mode);
/* Persistent CTAs of the tensor-core matcher (one per SM by default; 0 restores that).  An even budget runs the distance kernel
 * as CTA pairs (cta_group::2 MMAs, two SMs share every train tile), an odd one as single CTAs; env SBA_TC_PAIR=0 forces single CTAs.  With several pairs in flight on
 * one GPU, half the SMs per match lets two matches run side by side on longer spans. */
int sba_ctx_set_matcher_ctas(sba_ctx* ctx, int n_ctas);
/* Programmatic dependent launch along the kernel chain of a match / a pair (prep -> distance kernel -> re-rank -> fallback ->
 * finalize -> pair solve): each kernel is scheduled while its predecessor drains and waits on the device before touching
 * memory.  A latency knob: one C2 pair at a time 254 -> 240 us on B200, but several pairs in flight on one GPU lose ~3 % of
 * their throughput (waiting CTAs keep other pairs' kernels off their SMs), so it is off by default (env SBA_PDL=1 turns it
 * on for every new context).  Results are identical either way.  The three small kernels at the end of the chain (fallback,
 * finalize, pair solve) are always launched dependent: their early CTAs occupy next to nothing. */
int sba_ctx_set_dependent_launch(sba_ctx* ctx, int enable);
int sba_remap_plan_info(sba_ctx* ctx, int w, int h, int cube_size, int* tiled_available, int* tiled_preferred, int* n_tiles,
                        int* n_fallback_tiles, float trial_ms[4]);
/* The source-ordered form of a cube plan: available, which kernel (1 direct, 2 tiled, 3 source-ordered) the trials picked
 * for small (2 frames) and large batches, tile count, tiles whose source span does not fit the entry format (they read
 * the ordinary table inside the same kernel), trial_ms = {small, large} of the source-ordered kernel. */
int sba_remap_plan_sorted_info(sba_ctx* ctx, int w, int h, int cube_size, int* available, int* best_small, int* best_large,
                               int* n_tiles, int* n_fallback_tiles, float trial_ms[2]);
int sba_equi2cube_lut(sba_ctx* ctx, int w, int h, int cube_size, int32_t* lut_out, int mem);

/* equi2cube_surf::cube2equi_pixel for n keypoints (equi2cube_surf.cpp:19-76).
 * xy_in / xy_out: n interleaved (x, y) float pairs (cv::Point2f). */
int sba_cube2equi_points(sba_ctx* ctx, const float* xy_in, int n, int cube_size, int w, int h, float* xy_out, int mem);

/* ERP pixel -> unit bearing (spherical_bundle_adjuster.cpp:271-298), fp64 math.
 * bearings_out: n x float4 (x, y, z, 0) -- the layout the BA kernels read.
 * bearings64_out (optional, may be NULL): n x 3 doubles, the reference's vector<Point3d>. */
int sba_pixels_to_bearings(sba_ctx* ctx, const float* xy, int n, int w, int h, float* bearings_out, double* bearings64_out, int mem);

/* ---------------------------------------------------------------------------------------------
 * Matcher  (replaces feature_matcher::match_two_image, feature_matcher.cpp:42-59)
 *
 * Exact brute-force L2 kNN (k=2) of every query row over all train rows with OpenCV BFMatcher
 * semantics (fp32 distance in OpenCV's accumulation order, ties -> lower train index), then the
 * Lowe ratio test  d0 < ratio * d1  (feature_matcher.cpp:47,52).
 *
 * q [nq x dim], t [nt x dim] fp32 row-major.  dim must be a multiple of 16 (SURF: 64 or 128).
 * Survivors are written in ascending query order: query_idx/train_idx/dist [capacity nq].
 * knn_idx / knn_dist (optional, may be NULL): nq x 2, the raw kNN result; missing neighbours
 * (nt < 2) are -1 / +inf and such rows never pass the ratio test (the reference reads
 * knn[i][1] unconditionally there -- guarded here).
 * algo: SBA_MATCH_AUTO picks the tensor-core path when it applies.
 * ------------------------------------------------------------------------------------------- */
typedef enum sba_match_algo {
    SBA_MATCH_AUTO = 0,
    SBA_MATCH_SIMT_EXACT = 1, /* fp32 CUDA-core brute force in OpenCV's arithmetic order */
    SBA_MATCH_TENSOR = 2,     /* tcgen05 bf16x3 candidate filter + exact fp32 re-rank; what AUTO picks for large sets */
    SBA_MATCH_TENSOR_FP16 = 3 /* tcgen05 single-product fp16 candidate filter (a third of the tensor work, wider safety margin, more rows through the exact fallback) + exact fp32 re-rank */
} sba_match_algo;

int sba_knn2_ratio(sba_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim, float ratio,
                   int32_t* query_idx, int32_t* train_idx, float* dist, int32_t* n_matches,
                   int32_t* knn_idx, float* knn_dist, int mem, int algo);

/* Statistics of the last sba_knn2_ratio call on this context (tensor path diagnostics). */
typedef struct sba_match_stats {
    int algo_used;          /* sba_match_algo actually run */
    int n_fallback_rows;    /* query rows whose candidate set failed the error-bound test and were
                               re-scanned exactly */
    int n_tiles;            /* tensor-core tiles issued */
    int n_ctas;
    float max_rel_err;      /* largest observed |approx - exact| / (|a|^2 + max|b|^2) over the candidate
                               chunks (must stay below the bound coefficient 4e-5 the re-rank assumes) */
} sba_match_stats;
int sba_match_last_stats(sba_ctx* ctx, sba_match_stats* out);

/* A descriptor set that takes part in many matches (every frame of a sequence in all-pairs matching, the train set of
 * a sharded sweep) can be handed over once: the fp32 rows are copied to the device and, for 64-d rows, split into the
 * bf16 form the tensor-core kernel reads, so that step is not repeated per match.  The set is immutable and may be
 * used by any context of the same device.  sba_knn2_ratio_prepared is sba_knn2_ratio on two such sets (outputs and
 * `mem` as there; identical results). */
typedef struct sba_descriptors sba_descriptors;
int sba_descriptors_create(sba_ctx* ctx, const float* desc, int n, int dim, int mem, sba_descriptors** out);
int sba_descriptors_destroy(sba_descriptors* d);
int sba_descriptors_count(const sba_descriptors* d);
int sba_knn2_ratio_prepared(sba_ctx* ctx, const sba_descriptors* query, const sba_descriptors* train, float ratio,
                            int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out, int32_t* n_matches_out,
                            int32_t* knn_idx_out, float* knn_dist_out, int mem, int algo);

/* Gather matched keypoints (equi2cube_surf.cpp:107-113): out_left[i] = key_left[query_idx[i]],
 * out_right[i] = key_right[train_idx[i]] on (x, y) float pairs. */
int sba_gather_matches(sba_ctx* ctx, const float* key_left_xy, const float* key_right_xy, const int32_t* query_idx,
                       const int32_t* train_idx, int n_matches, float* out_left_xy, float* out_right_xy, int mem);

/* ---------------------------------------------------------------------------------------------
 * spherical_surf front-end geometry (spherical_surf.cpp:17-123; the crops of do_all :137-153)
 *
 *   sba_eular2rot           eular2rot (:17-45): float Euler angles -> 3x3 row-major doubles, host arithmetic
 *                           identical to the reference's (float-precision factors, (Rz Ry) Rx in double).
 *   sba_crop_rotated_lut    source index row*w+col of every pixel of the (h/4) x w band for one pitch, -1 where
 *                           the bounds check of :100 fails; bit-identical to the reference's index arithmetic.
 *   sba_crop_rotated_image  crop_rotated_image (:79-109) for n_images frames: out [n][h/4][w][3]; pixels the
 *                           reference leaves unwritten are 0.
 *   sba_spherical_crops     the four bands do_all cuts from an image (pitch 45, the plain band im(roi),
 *                           pitch -45, pitch -90) in ONE gather: out [n][4][h/4][w][3].
 *   sba_rotate_pixels       rotate_pixel (:48-77) on n (row, col) pairs with eular2rot(0, RAD(pitch), 0).
 *   sba_rotate_pixels_mat   the same with an arbitrary 3x3 rotation (row-major doubles), as rotate_pixel's signature allows.
 *   sba_rotate_keypoints    rotate_keypoint (:111-123): n (x, y) keypoints in band coordinates, in place.
 * ------------------------------------------------------------------------------------------- */
int sba_eular2rot(const float theta[3], double R_out[9]);
int sba_crop_rotated_lut(sba_ctx* ctx, int w, int h, float pitch_deg, int32_t* lut_out, int* n_patched, int mem);
int sba_crop_rotated_image(sba_ctx* ctx, const uint8_t* erp, int w, int h, int n_images, float pitch_deg, uint8_t* out, int mem);
int sba_spherical_crops(sba_ctx* ctx, const uint8_t* erp, int w, int h, int n_images, uint8_t* out, int mem);
int sba_rotate_pixels(sba_ctx* ctx, const int32_t* rc_in, int n, float pitch_deg, int w, int h, int32_t* rc_out, int mem);
int sba_rotate_pixels_mat(sba_ctx* ctx, const int32_t* rc_in, int n, const double R[9], int w, int h, int32_t* rc_out, int mem);
int sba_rotate_keypoints(sba_ctx* ctx, float* xy_inout, int n, float pitch_inv_deg, int w, int h, int mem);

/* ---------------------------------------------------------------------------------------------
 * Initial guess of the bundle adjuster (spherical_bundle_adjuster.cpp:47-181)
 *
 * b1, b2: n x 3 doubles (unit bearings, key_point_*_rect); idx: [n_samples x sample_n] row indices of the
 * random subsets (the reference: 80 subsets of n/4 matches drawn with std::random_shuffle, :126-137) -- all
 * three live where `mem` says; every output is a host pointer.
 *   sba_eight_point_null         per subset the null direction e [9] of its n x 9 epipolar system (the last row
 *                                of vt in :69, sign arbitrary) and/or the packed upper triangle of A^T A [45].
 *   sba_essential_to_candidates  :71-114 for one e: Euler angles of the two rotation candidates, unit translation,
 *                                validity flags (host arithmetic only).  {R1, R2} does not depend on the sign of e or
 *                                on SVD conventions; their order and the sign of T do (as with OpenCV).
 *   sba_initial_guess            the whole of :117-181: candidates of every subset, then the vote.
 * ------------------------------------------------------------------------------------------- */
int sba_eight_point_null(sba_ctx* ctx, const double* b1, const double* b2, int n, const int32_t* idx, int n_samples, int sample_n,
                         double* ata_out, double* e_out, int mem);
int sba_essential_to_candidates(const double e[9], float R1_vec[3], float R2_vec[3], float T_vec[3], int* R1_valid, int* R2_valid);
int sba_initial_guess(sba_ctx* ctx, const double* b1, const double* b2, int n, const int32_t* idx, int n_samples, int sample_n,
                      float R_vec_out[3], float T_vec_out[3], int* n_candidates, int mem);

/* ---------------------------------------------------------------------------------------------
 * Rotation-only bundle adjustment
 * (replaces ba_spherical_costfunctor_rot_only + ceres::Solve, spherical_bundle_adjuster.cpp:892-945,
 *  :183-217, :334-338)
 *
 * A problem owns the observations on the device: b1, b2 are n_obs x float4 unit bearings
 * (x, y, z, unused); cam (may be NULL == all zero) gives the rotation block each observation
 * belongs to, 0 <= cam < n_cam.  The reference has exactly one block (:943); n_cam > 1 is the
 * multi-camera extension of BASELINE config 4.  Observations are grouped by camera at creation.
 *
 * residual_i = d2*b2_i - (R(r_cam) * d1*b1_i - t)            (:896-916)
 * loss       = Huber(huber_delta) on |residual_i|^2, huber_delta <= 0 means no loss (:943)
 * ------------------------------------------------------------------------------------------- */
typedef struct sba_ba_problem sba_ba_problem;

int sba_ba_problem_create(sba_ctx* ctx, const float* b1, const float* b2, const int32_t* cam, int64_t n_obs, int n_cam,
                          int mem, sba_ba_problem** out);
int sba_ba_problem_destroy(sba_ba_problem* p);

/* Multi-GPU: this rank holds a shard of the residuals.  After each evaluation the per-camera
 * normal-equation blocks are summed over ranks by `allreduce(buffer, count, user)`, which must
 * sum `count` doubles in place on the context's stream (bench/tests pass an NCCL all-reduce).
 * Pass NULL to return to single-GPU operation. */
typedef int (*sba_allreduce_fn)(void* device_buffer, int64_t count, void* user);
int sba_ba_problem_set_allreduce(sba_ba_problem* p, sba_allreduce_fn fn, void* user);

/* Peer-memory exchange (one process per GPU on one NVLink/NVSwitch box): instead of a host-launched
 * all-reduce between kernels, the evaluation kernel itself publishes this rank's [n_cam x 10] blocks in
 * a buffer every peer has mapped (CUDA IPC), waits for the peers' sequence flags over NVLink and sums all
 * ranks' blocks in rank order -- identical bits on every rank, one kernel launch per LM iteration.
 *   1. every rank: sba_comm_create(ctx, rank, world, max_cameras, &comm, handle)   (handle: 64 bytes)
 *   2. ranks exchange the handles out of band (torch.distributed all_gather in the harness)
 *   3. every rank: sba_comm_connect(comm, all_handles)     (world x 64 bytes, in rank order)
 *   4. sba_ba_problem_set_comm(problem, comm) on the problem holding this rank's residual shard.
 * All ranks must then issue the same sequence of solves/evaluations on their problems. */
typedef struct sba_comm sba_comm;
#define SBA_COMM_HANDLE_BYTES 64
int sba_comm_create(sba_ctx* ctx, int rank, int world, int max_cameras, sba_comm** out, void* ipc_handle_out);
int sba_comm_connect(sba_comm* comm, const void* ipc_handles);
int sba_comm_destroy(sba_comm* comm);
int sba_ba_problem_set_comm(sba_ba_problem* p, sba_comm* comm);

/* One evaluation at rotations r [n_cam x 3] (axis-angle, host pointer always).
 * Outputs, each optional (NULL to skip), located per `mem`:
 *   res  n_obs x 3 fp32 and jac n_obs x 9 fp32 (row-major d res_a / d r_k): the RAW functor values
 *        in the caller's original observation order;
 *   H n_cam x 6 (xx,xy,xz,yy,yz,zz), g n_cam x 3, cost n_cam, fp64: Huber-corrected J^T J, J^T r
 *        and 1/2 sum rho per camera. */
int sba_ba_rot_eval(sba_ba_problem* p, const double* r, const double t[3], double d1, double d2, double huber_delta,
                    float* res, float* jac, double* H, double* g, double* cost, int mem);

typedef struct sba_solve_summary {
    int iterations;      /* LM iterations run (successful + unsuccessful + invalid) */
    int num_successful;
    int termination;     /* 0 max iterations, 1 function tol, 2 gradient tol, 3 parameter tol, 4 failure (five invalid steps in a row),
                            5 minimum trust-region radius reached (Ceres reports CONVERGENCE for it) */
    int evaluations;     /* residual+Jacobian passes over the observations */
    double initial_cost;
    double final_cost;
    double final_radius;
} sba_solve_summary;

/* Levenberg-Marquardt with Ceres' default trust-region policy (one radius for the whole problem),
 * at most max_iter iterations (the reference sets 50, :336).  r_inout [n_cam x 3] host pointer. */
int sba_ba_rot_solve(sba_ba_problem* p, double* r_inout, const double t[3], double d1, double d2, double huber_delta,
                     int max_iter, sba_solve_summary* summary);

/* Translation-only variant (ba_spherical_costfunctor_tran_only, spherical_bundle_adjuster.cpp:948-1002,
 * solved at :208-209): same residual, the rotations r_fixed [n_cam x 3] are constants and the free block
 * is the translation t [n_cam x 3] (the reference has one camera: a single 3-vector); d res / d t = +I.
 * Outputs of the evaluation as in sba_ba_rot_eval (the Jacobian is the identity and is not materialised). */
int sba_ba_tran_eval(sba_ba_problem* p, const double* r_fixed, const double* t, double d1, double d2, double huber_delta,
                     float* res, double* H, double* g, double* cost, int mem);
int sba_ba_tran_solve(sba_ba_problem* p, const double* r_fixed, double* t_inout, double d1, double d2, double huber_delta,
                      int max_iter, sba_solve_summary* summary);

/* Depth-only block (ba_spherical_costfunctor_d_only, spherical_bundle_adjuster.cpp:1005-1063): per match
 * the free block is its own depth pair d[i] = (d1, d2) >= 0; residuals d2*b2 - (R(r)(d1*b1) - t) and the two
 * barrier terms lambda*exp(-c*d); no loss function.  One camera pair only (n_cam == 1).
 *   sba_ba_d_eval   raw functor values: res [n x 5], jac [n x 10] (row-major 5 x 2), cost [n] = 1/2 |res|^2.
 *   sba_ba_d_solve  the first ceres::Solve of solve_problem (:196-197): LM over all blocks as ONE problem,
 *                   bounds handled like Ceres (projection + projected Armijo line search).  d_inout [n x 2]
 *                   lives where `mem` says; line_search_trials (optional) counts trial points beyond alpha = 1. */
int sba_ba_d_eval(sba_ba_problem* p, const double r[3], const double t[3], const double* d, double lambda, double c,
                  double* res, double* jac, double* cost, int mem);
int sba_ba_d_solve(sba_ba_problem* p, const double r[3], const double t[3], double* d_inout, double lambda, double c,
                   int max_iter, sba_solve_summary* summary, int* line_search_trials, int mem);

/* spherical_bundle_adjuster::solve_problem (spherical_bundle_adjuster.cpp:183-217): depth stage, rotation
 * stage, translation stage, each a full LM solve starting from the previous stage's result.  r_inout,
 * t_inout: host 3-vectors (init_rot, init_tran); d_inout [n x 2] (init_d) where `mem` says;
 * summaries: optional array of three (depth, rotation, translation). */
int sba_ba_solve_problem(sba_ba_problem* p, double r_inout[3], double t_inout[3], double* d_inout, double huber_delta,
                         int max_iter, sba_solve_summary summaries[3], int mem);

/* Device-timed evaluation loop for benchmarking: runs `iters` fused evaluations (residual +
 * Jacobian + per-camera normal equations) back to back at r and returns the mean kernel time. */
int sba_ba_rot_eval_timed(sba_ba_problem* p, const double* r, const double t[3], double d1, double d2, double huber_delta,
                          int materialise, int iters, float* mean_ms);

/* ---------------------------------------------------------------------------------------------
 * The whole hot path for one ERP pair in one call
 * (equi2cube_surf::do_all, equi2cube_surf.cpp:78-122, minus SURF; then the bearing conversion and the
 *  rotation-only solve of spherical_bundle_adjuster::do_bundle_adjustment, :268-298 and :202-203).
 *
 *   erp_left/right   h x w x 3 images, or both NULL to skip the remap stage.
 *   strip_*_out      optional cube strips (cube_size x 6*cube_size x 3); NULL keeps them on the device.
 *   desc_*, key_*    SURF descriptors [n x dim] and keypoints [n x 2] in cube-strip coordinates
 *                    (what detect/compute on the strips return; SURF itself stays with the caller).
 *   query/train_idx_out, dist_out  optional match list (capacity n_left).
 *   r0, t, d1, d2, huber_delta, max_iter  as in sba_ba_rot_solve.
 * All data pointers live where `mem` says; `result` is a host struct.  With SBA_MEM_HOST the call
 * copies inputs in, runs on the device and copies the requested outputs back before returning.
 * ------------------------------------------------------------------------------------------- */
typedef struct sba_pair_result {
    double rotation[3];
    int n_matches;
    int lm_iterations;
    int lm_termination;
    int reserved;
    double initial_cost;
    double final_cost;
} sba_pair_result;

int sba_pair_rotation(sba_ctx* ctx, const uint8_t* erp_left, const uint8_t* erp_right, int w, int h, int cube_size,
                      uint8_t* strip_left_out, uint8_t* strip_right_out, const float* desc_left, int n_left,
                      const float* desc_right, int n_right, int dim, const float* key_left_xy, const float* key_right_xy,
                      float ratio, const double r0[3], const double t[3], double d1, double d2, double huber_delta,
                      int max_iter, int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out,
                      sba_pair_result* result, int mem);

/* The same call split in two so that a host thread can keep several pairs in flight (one per context):
 * _begin queues all stream work of the pair and returns without waiting; _end waits for it, finishes the
 * solve if its first batch of evaluations was not enough, copies the requested outputs back and fills
 * `result`.  All buffers passed to _begin must stay valid until _end returns.  One call in flight per
 * context (the staging buffers belong to the context): pipelining N pairs takes N contexts on N streams. */
typedef struct sba_pair_call sba_pair_call;
int sba_pair_rotation_begin(sba_ctx* ctx, const uint8_t* erp_left, const uint8_t* erp_right, int w, int h, int cube_size,
                            uint8_t* strip_left_out, uint8_t* strip_right_out, const float* desc_left, int n_left,
                            const float* desc_right, int n_right, int dim, const float* key_left_xy, const float* key_right_xy,
                            float ratio, const double r0[3], const double t[3], double d1, double d2, double huber_delta,
                            int max_iter, int32_t* query_idx_out, int32_t* train_idx_out, float* dist_out, int mem,
                            sba_pair_call** call_out);
int sba_pair_rotation_end(sba_pair_call* call, sba_pair_result* result);

#ifdef __cplusplus
}
#endif
#endif /* SBA_B200_H */
