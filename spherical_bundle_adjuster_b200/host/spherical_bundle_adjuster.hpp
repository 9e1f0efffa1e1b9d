// Drop-in for the rotation-only path of the reference's spherical_bundle_adjuster.hpp.
//
// The class keeps the reference's constructor / set_omp / do_bundle_adjustment signatures
// (spherical_bundle_adjuster.hpp:15-23).  The Ceres functor struct keeps its name and its add_residual
// argument list minus the ceres::Problem (Ceres is not a dependency any more): `solve` does what
// `add_residual(problem, ...)` followed by `ceres::Solve(opt, &problem, &summary)` did
// (spherical_bundle_adjuster.cpp:202-203), on the GPU.
#pragma once
#include <algorithm>
#include <array>
#include <numeric>
#include <vector>

#include "equi2cube_surf.hpp"
#include "spherical_surf.hpp"
#include "sba_b200.h"

struct ba_spherical_costfunctor_rot_only
{
    // Residual + Jacobian evaluation and LM solve of spherical_bundle_adjuster.cpp:892-945 on the GPU.
    // init_rot is updated in place (the single shared 3-vector parameter block, :943); t = init_tran,
    // d1 = init_d[0][0], d2 = init_d[1][0] for every residual exactly as the reference passes them
    // (:938-942); Huber(1.0); at most max_num_iterations LM iterations (the reference sets 50, :336).
    static sba_solve_summary solve(std::vector<cv::Point3d>& key_point_left_rect
                                 , std::vector<cv::Point3d>& key_point_right_rect
                                 , double* init_rot
                                 , double* init_tran
                                 , std::vector<std::array<double, 2>>& init_d
                                 , int match_num
                                 , int max_num_iterations = 50);
};

// Depth-only block (spherical_bundle_adjuster.cpp:1005-1063): every match owns its depth pair init_d[i],
// bounded below by 0, no loss function, lambda = c = 1.  init_d is updated in place.
struct ba_spherical_costfunctor_d_only
{
    static sba_solve_summary solve(std::vector<cv::Point3d>& key_point_left_rect
                                 , std::vector<cv::Point3d>& key_point_right_rect
                                 , double* init_rot
                                 , double* init_tran
                                 , std::vector<std::array<double, 2>>& init_d
                                 , int match_num
                                 , int max_num_iterations = 50);
};

// Translation-only block (spherical_bundle_adjuster.cpp:948-1002): init_tran is updated in place.
struct ba_spherical_costfunctor_tran_only
{
    static sba_solve_summary solve(std::vector<cv::Point3d>& key_point_left_rect
                                 , std::vector<cv::Point3d>& key_point_right_rect
                                 , double* init_rot
                                 , double* init_tran
                                 , std::vector<std::array<double, 2>>& init_d
                                 , int match_num
                                 , int max_num_iterations = 50);
};

// The three fields of ceres::Solver::Options the reference sets (spherical_bundle_adjuster.cpp:334-338);
// linear_solver_type is not needed: every stage's reduced system is empty, the exact block solve IS the
// ITERATIVE_SCHUR answer.
struct sba_solver_options
{
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = true;
    int num_threads = 1;
};

class spherical_bundle_adjuster
{
    public:
    spherical_bundle_adjuster(double roll = 0, double pitch = 0, double yaw = 0, double tx = 0, double ty = 0, double tz = 0, double d = 0)
    : expected_roll(roll), expected_pitch(pitch), expected_yaw(yaw), expected_tx(tx), expected_ty(ty), expected_tz(tz), expected_d(d) {}
    ~spherical_bundle_adjuster() {}

    void set_omp(int num_proc);
    // spherical_surf front-end (:252-266) -> bearings (:268-298) -> initial_guess (:302-306) -> solve_problem
    // (:333-345), the reference's call chain.  Needs SURF (real OpenCV) for the front-end.
    void do_bundle_adjustment(const cv::Mat &im_left, const cv::Mat &im_right);

    // Everything of do_bundle_adjustment after the front-end: matched ERP keypoints -> bearings -> initial guess
    // -> three-stage solve.  init_rot/init_tran: NULL = run initial_guess like the reference (negated Euler angles
    // fed as the angle-axis start, :330-331); otherwise start from the given values.
    // Results in result_rot / result_tran / result_d.
    void adjust(const std::vector<cv::KeyPoint>& left_key, const std::vector<cv::KeyPoint>& right_key, int im_width, int im_height,
                const double* init_rot = nullptr, const double* init_tran = nullptr);

    // private in the reference (spherical_bundle_adjuster.hpp:26-38)
    void eight_point_estimation(int im_width, int im_height
                                , std::vector<cv::Point3d>& key_point_left_rect, std::vector<cv::Point3d>& key_point_right_rect
                                , cv::Vec3f& R1_vec, cv::Vec3f& R2_vec, cv::Vec3f& T_vec
                                , bool& R1_valid, bool& R2_valid
                                , int match_size);
    void initial_guess(int im_width, int im_height
                        , std::vector<cv::Point3d>& key_point_left_rect, std::vector<cv::Point3d>& key_point_right_rect
                        , cv::Vec3f& R_vec_out, cv::Vec3f& T_vec_out
                        , int match_size);

    // The post-SURF part: matched ERP keypoints -> bearings (spherical_bundle_adjuster.cpp:268-298) ->
    // rotation-only solve.  Returns the rotation vector in rot[3].
    sba_solve_summary adjust_rotation(const std::vector<cv::KeyPoint>& left_key, const std::vector<cv::KeyPoint>& right_key, int im_width,
                                      int im_height, double rot[3]);

    // The reference's private solve_problem (spherical_bundle_adjuster.cpp:183-217), public here so callers
    // that have their own front-end can use it: depth stage -> rotation stage -> translation stage, all three
    // parameter sets updated in place; prints the three brief reports when the options ask for progress.
    void solve_problem(sba_solver_options& opt
                    , std::vector<cv::Point3d>& key_point_left_rect
                    , std::vector<cv::Point3d>& key_point_right_rect
                    , double* init_rot
                    , double* init_tran
                    , std::vector<std::array<double, 2>>& init_d
                    , int match_num);
    sba_solve_summary stage_summaries[3] = {};   // depth, rotation, translation of the last solve_problem

    double result_rot[3] = {0, 0, 0};
    double result_tran[3] = {0, 0, 0};
    std::vector<std::array<double, 2>> result_d;

    private:
    double expected_roll, expected_pitch, expected_yaw, expected_tx, expected_ty, expected_tz, expected_d;
    int num_proc = 1;
};

// spherical_bundle_adjuster.hpp:182-210 of the reference: a shuffled 0..size-1 handed out one by one.  Same
// libstdc++ call (std::random_shuffle on std::rand), so the subsets are the ones the reference would draw.
class random_array
{
    public:
    random_array(int size) : rand_arr(size)
    {
        size_ = size;
        count_ = 0;
        std::iota(rand_arr.begin(), rand_arr.end(), 0);
        std::random_shuffle(rand_arr.begin(), rand_arr.end());
    }
    int get_rand()
    {
        int retval = rand_arr[count_];
        count_ = (count_ + 1) % size_;
        return retval;
    }

    private:
    int size_;
    std::vector<int> rand_arr;
    int count_;
};
