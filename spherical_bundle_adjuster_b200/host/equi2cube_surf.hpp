// Drop-in for the reference's equi2cube_surf.hpp:7-18.
#pragma once
#include "equi2cube.hpp"
#include "feature_matcher.hpp"

class equi2cube_surf
{
    public:
    void set_omp(int num_proc);
    void set_cube_size(int cube_size);
    void cube2equi_pixel(cv::Point2f& cube_pixel, cv::Point2f& equi_pixel, int cube_size, int im_width, int im_height);
    void do_all(const cv::Mat &im_left, const cv::Mat &im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num);

    // Not in the reference: the part of do_all after SURF (match, cube->ERP, gather), for callers that
    // already hold keypoints/descriptors of the two cube strips.
    void match_and_lift(const std::vector<cv::KeyPoint>& key_left_cube, const std::vector<cv::KeyPoint>& key_right_cube, const cv::Mat& desc_left,
                        const cv::Mat& desc_right, int im_width, int im_height, std::vector<cv::KeyPoint>& left_key,
                        std::vector<cv::KeyPoint>& right_key, std::vector<cv::DMatch>& matches);

    private:
    int num_proc = 1;
    int cube_size = 600;   // the reference leaves this unset until set_cube_size; its test uses 600
};
