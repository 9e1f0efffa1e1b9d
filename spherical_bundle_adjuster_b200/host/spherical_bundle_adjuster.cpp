// Facade: rotation-only spherical bundle adjustment.  Replaces spherical_bundle_adjuster.cpp:268-298
// (pixel -> bearing), :892-945 (functor + add_residual) and the ceres::Solve call at :203.
#include "spherical_bundle_adjuster.hpp"

#include "sba_host_ctx.hpp"

#include <cmath>
#include <cstdint>
#include <cstdio>

namespace {

// unit vectors as the float4 records the device problem reads
sba_ba_problem* make_problem(std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right, int match_num)
{
    std::vector<float> b1(4 * (size_t)match_num), b2(4 * (size_t)match_num);
    for (int i = 0; i < match_num; i++) {
        b1[4 * i] = (float)left[i].x; b1[4 * i + 1] = (float)left[i].y; b1[4 * i + 2] = (float)left[i].z; b1[4 * i + 3] = 0.f;
        b2[4 * i] = (float)right[i].x; b2[4 * i + 1] = (float)right[i].y; b2[4 * i + 2] = (float)right[i].z; b2[4 * i + 3] = 0.f;
    }
    sba_ba_problem* prob = nullptr;
    sba_host::check(sba_ba_problem_create(sba_host::ctx(), b1.data(), b2.data(), nullptr, match_num, 1, SBA_MEM_HOST, &prob));
    return prob;
}

const char* termination_name(int t)
{
    switch (t) {
        case 0: return "NO_CONVERGENCE";
        case 4: return "FAILURE";
        default: return "CONVERGENCE";
    }
}

}  // namespace

sba_solve_summary ba_spherical_costfunctor_d_only::solve(std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right, double* init_rot,
                                                         double* init_tran, std::vector<std::array<double, 2>>& init_d, int match_num,
                                                         int max_num_iterations)
{
    sba_solve_summary sum{};
    if (match_num <= 0) return sum;
    sba_ba_problem* prob = make_problem(left, right, match_num);
    // std::array<double, 2> is two contiguous doubles: init_d.data() is the [n x 2] table
    int status = sba_ba_d_solve(prob, init_rot, init_tran, init_d[0].data(), 1.0, 1.0 /* lambda, c: :1058-1059 */, max_num_iterations, &sum,
                                nullptr, SBA_MEM_HOST);
    sba_ba_problem_destroy(prob);
    sba_host::check(status);
    return sum;
}

sba_solve_summary ba_spherical_costfunctor_tran_only::solve(std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right, double* init_rot,
                                                            double* init_tran, std::vector<std::array<double, 2>>& init_d, int match_num,
                                                            int max_num_iterations)
{
    sba_solve_summary sum{};
    if (match_num <= 0) return sum;
    const double d1 = init_d.size() > 0 ? init_d[0][0] : 1.0, d2 = init_d.size() > 1 ? init_d[1][0] : d1;   // :998-999
    sba_ba_problem* prob = make_problem(left, right, match_num);
    int status = sba_ba_tran_solve(prob, init_rot, init_tran, d1, d2, 1.0, max_num_iterations, &sum);
    sba_ba_problem_destroy(prob);
    sba_host::check(status);
    return sum;
}

void spherical_bundle_adjuster::solve_problem(sba_solver_options& opt, std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right,
                                              double* init_rot, double* init_tran, std::vector<std::array<double, 2>>& init_d, int match_num)
{
    for (auto& s : stage_summaries) s = sba_solve_summary{};
    if (match_num < 2) return;
    sba_ba_problem* prob = make_problem(left, right, match_num);   // one upload for the three stages
    int status = sba_ba_solve_problem(prob, init_rot, init_tran, init_d[0].data(), 1.0, opt.max_num_iterations, stage_summaries, SBA_MEM_HOST);
    sba_ba_problem_destroy(prob);
    sba_host::check(status);
    if (opt.minimizer_progress_to_stdout) {
        for (const auto& s : stage_summaries)   // the shape of ceres' Summary::BriefReport (:198, :204, :210)
            std::printf("Ceres Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s\n", s.iterations,
                        s.initial_cost, s.final_cost, termination_name(s.termination));
        std::printf("rotation vector in degree %g %g %g\n", init_rot[0] / M_PI * 180.0, init_rot[1] / M_PI * 180.0, init_rot[2] / M_PI * 180.0);
        std::printf("translation vector %g %g %g\n", init_tran[0], init_tran[1], init_tran[2]);
    }
}

sba_solve_summary ba_spherical_costfunctor_rot_only::solve(std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right, double* init_rot,
                                                           double* init_tran, std::vector<std::array<double, 2>>& init_d, int match_num,
                                                           int max_num_iterations)
{
    sba_solve_summary sum{};
    if (match_num <= 0) return sum;
    // the reference hands init_d[0][0] and init_d[1][0] to EVERY residual (:941-942); preserved
    const double d1 = init_d.size() > 0 ? init_d[0][0] : 1.0, d2 = init_d.size() > 1 ? init_d[1][0] : d1;
    sba_ba_problem* prob = make_problem(left, right, match_num);
    int status = sba_ba_rot_solve(prob, init_rot, init_tran, d1, d2, 1.0 /* HuberLoss(1.0), :943 */, max_num_iterations, &sum);
    sba_ba_problem_destroy(prob);
    sba_host::check(status);
    return sum;
}

void spherical_bundle_adjuster::set_omp(int num_proc) { this->num_proc = num_proc; }

sba_solve_summary spherical_bundle_adjuster::adjust_rotation(const std::vector<cv::KeyPoint>& left_key, const std::vector<cv::KeyPoint>& right_key,
                                                             int im_width, int im_height, double rot[3])
{
    const int n = (int)left_key.size();
    sba_solve_summary sum{};
    if (n == 0) return sum;
    // pixel -> bearing for both sides in one 2n batch (spherical_bundle_adjuster.cpp:271-298)
    std::vector<float> px(4 * (size_t)n), b(8 * (size_t)n);
    for (int i = 0; i < n; i++) {
        px[2 * i] = left_key[i].pt.x; px[2 * i + 1] = left_key[i].pt.y;
        px[2 * (n + i)] = right_key[i].pt.x; px[2 * (n + i) + 1] = right_key[i].pt.y;
    }
    sba_host::check(sba_pixels_to_bearings(sba_host::ctx(), px.data(), 2 * n, im_width, im_height, b.data(), nullptr, SBA_MEM_HOST));
    sba_ba_problem* prob = nullptr;
    sba_host::check(sba_ba_problem_create(sba_host::ctx(), b.data(), b.data() + 4 * (size_t)n, nullptr, n, 1, SBA_MEM_HOST, &prob));
    const double t[3] = {expected_tx, expected_ty, expected_tz};
    // d = expected_d for both cameras like init_d (:325-326); 0 would zero every residual, so the usual
    // unit sphere is used when the caller left it at the default
    const double d = expected_d != 0.0 ? expected_d : 1.0;
    int status = sba_ba_rot_solve(prob, rot, t, d, d, 1.0, 50, &sum);
    sba_ba_problem_destroy(prob);
    sba_host::check(status);
    return sum;
}

static void flatten(const std::vector<cv::Point3d>& pts, int n, std::vector<double>& out)
{
    out.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++) { out[3 * i] = pts[i].x; out[3 * i + 1] = pts[i].y; out[3 * i + 2] = pts[i].z; }
}

void spherical_bundle_adjuster::eight_point_estimation(int, int, std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right, cv::Vec3f& R1_vec,
                                                       cv::Vec3f& R2_vec, cv::Vec3f& T_vec, bool& R1_valid, bool& R2_valid, int match_size)
{
    std::vector<double> b1, b2;
    flatten(left, match_size, b1);
    flatten(right, match_size, b2);
    std::vector<int32_t> idx(match_size);
    std::iota(idx.begin(), idx.end(), 0);
    double e[9];
    sba_host::check(sba_eight_point_null(sba_host::ctx(), b1.data(), b2.data(), match_size, idx.data(), 1, match_size, nullptr, e, SBA_MEM_HOST));
    int v1 = 0, v2 = 0;
    sba_host::check(sba_essential_to_candidates(e, R1_vec.val, R2_vec.val, T_vec.val, &v1, &v2));
    R1_valid = v1 != 0;
    R2_valid = v2 != 0;
}

void spherical_bundle_adjuster::initial_guess(int, int, std::vector<cv::Point3d>& left, std::vector<cv::Point3d>& right, cv::Vec3f& R_vec_out,
                                              cv::Vec3f& T_vec_out, int match_size)
{
    // :124-137: 80 times a fresh shuffle, the first quarter of it is the subset
    const int n_samples = 80, sample_n = (int)(match_size * 0.25);
    std::vector<int32_t> idx((size_t)n_samples * sample_n);
    for (int s = 0; s < n_samples; s++) {
        random_array rand_arr(match_size);
        for (int i = 0; i < sample_n; i++) idx[(size_t)s * sample_n + i] = rand_arr.get_rand();
    }
    std::vector<double> b1, b2;
    flatten(left, match_size, b1);
    flatten(right, match_size, b2);
    sba_host::check(sba_initial_guess(sba_host::ctx(), b1.data(), b2.data(), match_size, idx.data(), n_samples, sample_n, R_vec_out.val,
                                      T_vec_out.val, nullptr, SBA_MEM_HOST));
}

void spherical_bundle_adjuster::adjust(const std::vector<cv::KeyPoint>& left_key, const std::vector<cv::KeyPoint>& right_key, int im_width,
                                       int im_height, const double* init_rot, const double* init_tran)
{
    const int n = (int)left_key.size();
    for (int k = 0; k < 3; k++) { result_rot[k] = init_rot ? init_rot[k] : 0.0; result_tran[k] = init_tran ? init_tran[k] : 0.0; }
    result_d.assign(n, std::array<double, 2>{{expected_d, expected_d}});   // :325-326
    if (n < 8) return;
    // pixel -> radian -> unit vector in double (:268-298), both sides in one batch
    std::vector<float> px(4 * (size_t)n);
    std::vector<double> b(6 * (size_t)n);
    for (int i = 0; i < n; i++) {
        px[2 * i] = left_key[i].pt.x; px[2 * i + 1] = left_key[i].pt.y;
        px[2 * (n + i)] = right_key[i].pt.x; px[2 * (n + i) + 1] = right_key[i].pt.y;
    }
    sba_host::check(sba_pixels_to_bearings(sba_host::ctx(), px.data(), 2 * n, im_width, im_height, nullptr, b.data(), SBA_MEM_HOST));
    std::vector<cv::Point3d> L(n), R(n);
    for (int i = 0; i < n; i++) {
        L[i] = cv::Point3d(b[3 * i], b[3 * i + 1], b[3 * i + 2]);
        R[i] = cv::Point3d(b[3 * (n + i)], b[3 * (n + i) + 1], b[3 * (n + i) + 2]);
    }
    if (!init_rot || !init_tran) {
        cv::Vec3f R_vec_out, T_vec_out;                                     // :302-306
        initial_guess(im_width, im_height, L, R, R_vec_out, T_vec_out, n);
        for (int k = 0; k < 3; k++) { result_rot[k] = -R_vec_out[k]; result_tran[k] = T_vec_out[k]; }   // :330-331
    }
    sba_solver_options options;              // :333-338
    options.max_num_iterations = 50;
    options.minimizer_progress_to_stdout = true;
    options.num_threads = num_proc;
    solve_problem(options, L, R, result_rot, result_tran, result_d, n);
}

void spherical_bundle_adjuster::do_bundle_adjustment(const cv::Mat& im_left, const cv::Mat& im_right)
{
    std::vector<cv::KeyPoint> left_key, right_key;
    int match_size = 0, total_key_num = 0;
    cv::Mat match_output;
    spherical_surf fm;                       // :263-266
    fm.set_omp(num_proc);
    fm.do_all(im_left, im_right, left_key, right_key, match_size, match_output, total_key_num);
    adjust(left_key, right_key, im_left.cols, im_left.rows);
}
