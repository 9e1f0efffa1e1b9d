// Drop-in for the reference's spherical_surf.hpp:9-30 -- the front-end do_bundle_adjustment calls
// (spherical_bundle_adjuster.cpp:264-266): four pitched/plain equatorial bands per image instead of a cubemap.
#pragma once
#define _USE_MATH_DEFINES
#include <cmath>

#include "feature_matcher.hpp"

#ifndef RAD
#define RAD(x) M_PI*(x)/180.0
#define DEGREE(x) 180.0*(x)/M_PI
#endif

class spherical_surf
{
    public:
    void set_omp(int num_proc);
    void do_all(const cv::Mat &im_left, const cv::Mat &im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num);

    cv::Mat eular2rot(cv::Vec3f theta);
    cv::Vec2i rotate_pixel(const cv::Vec2i& in_vec, cv::Mat& rot_mat, int width, int height);

    // private in the reference (spherical_surf.hpp:19-20); public here so callers with their own detector can use them
    cv::Mat crop_rotated_image(float pitch_rot, const cv::Mat& im);
    void rotate_keypoint(float pitch_rot_inv, std::vector<cv::KeyPoint>& key, int width, int height);

    // Not in the reference: the four bands of do_all (:137-143: pitch 45, plain band, -45, -90) from ONE gather.
    void crop_bands(const cv::Mat& im, cv::Mat bands[4]);
    // Not in the reference: everything of do_all after SURF (:180-232) -- band keypoints back to ERP coordinates,
    // concatenation in band order, matching, gathering the matched pairs.  key_*[b] / desc_*[b]: band b of one image.
    void lift_and_match(std::vector<cv::KeyPoint> key_left[4], std::vector<cv::KeyPoint> key_right[4], const cv::Mat desc_left[4],
                        const cv::Mat desc_right[4], int im_width, int im_height, std::vector<cv::KeyPoint>& left_key,
                        std::vector<cv::KeyPoint>& right_key, std::vector<cv::DMatch>& matches);

    private:
    int num_proc = 1;
};
