// Facade: cubemap front-end.  Replaces equi2cube_surf.cpp:19-122 of the reference.
#include "equi2cube_surf.hpp"

#include "sba_host_ctx.hpp"

void equi2cube_surf::set_omp(int num_proc) { this->num_proc = num_proc; }

void equi2cube_surf::set_cube_size(int cube_size) { this->cube_size = cube_size; }

// equi2cube_surf.cpp:19-76 for one keypoint (do_all batches all keypoints in one call instead)
void equi2cube_surf::cube2equi_pixel(cv::Point2f& cube_pixel, cv::Point2f& equi_pixel, int cube_size, int im_width, int im_height)
{
    float in[2] = {cube_pixel.x, cube_pixel.y}, out[2];
    sba_host::check(sba_cube2equi_points(sba_host::ctx(), in, 1, cube_size, im_width, im_height, out, SBA_MEM_HOST));
    equi_pixel.x = out[0];
    equi_pixel.y = out[1];
}

static void lift_keypoints(const std::vector<cv::KeyPoint>& cube, std::vector<cv::KeyPoint>& equi, int cube_size, int w, int h)
{
    const int n = (int)cube.size();
    std::vector<float> in(2 * (size_t)n), out(2 * (size_t)n);
    for (int i = 0; i < n; i++) { in[2 * i] = cube[i].pt.x; in[2 * i + 1] = cube[i].pt.y; }
    if (n) sba_host::check(sba_cube2equi_points(sba_host::ctx(), in.data(), n, cube_size, w, h, out.data(), SBA_MEM_HOST));
    equi = cube;   // every other KeyPoint field carries over (equi2cube_surf.cpp:100-101)
    for (int i = 0; i < n; i++) { equi[i].pt.x = out[2 * i]; equi[i].pt.y = out[2 * i + 1]; }
}

void equi2cube_surf::match_and_lift(const std::vector<cv::KeyPoint>& key_left_cube, const std::vector<cv::KeyPoint>& key_right_cube,
                                    const cv::Mat& desc_left, const cv::Mat& desc_right, int im_width, int im_height,
                                    std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, std::vector<cv::DMatch>& matches)
{
    feature_matcher fm;
    matches = fm.match_two_image(desc_left, desc_right);
    // The reference sizes BOTH lifted vectors by the LEFT keypoint count (equi2cube_surf.cpp:96-104) and
    // reads key_right_cube out of bounds when the right image has fewer keypoints; each side is lifted
    // over its own length here.
    std::vector<cv::KeyPoint> key_left_equi, key_right_equi;
    lift_keypoints(key_left_cube, key_left_equi, cube_size, im_width, im_height);
    lift_keypoints(key_right_cube, key_right_equi, cube_size, im_width, im_height);
    left_key.resize(matches.size());
    right_key.resize(matches.size());
    for (size_t i = 0; i < matches.size(); i++) {
        left_key[i] = key_left_equi[matches[i].queryIdx];     // equi2cube_surf.cpp:107-113
        right_key[i] = key_right_equi[matches[i].trainIdx];
    }
}

void equi2cube_surf::do_all(const cv::Mat& im_left, const cv::Mat& im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num)
{
    // equi2cube_surf.cpp:78-122
    int im_width = im_left.cols, im_height = im_left.rows;
    equi2cube cube;
    cv::Mat cubemap1 = cube.get_all(im_left, cube_size);
    cv::Mat cubemap2 = cube.get_all(im_right, cube_size);

    feature_matcher fm;
    std::vector<cv::KeyPoint> key_left_cube = fm.detect_key_point(cubemap1);
    std::vector<cv::KeyPoint> key_right_cube = fm.detect_key_point(cubemap2);
    cv::Mat desc_left = fm.comput_descriptor(cubemap1, key_left_cube);
    cv::Mat desc_right = fm.comput_descriptor(cubemap2, key_right_cube);

    std::vector<cv::DMatch> matches;
    match_and_lift(key_left_cube, key_right_cube, desc_left, desc_right, im_width, im_height, left_key, right_key, matches);
    match_output = fm.draw_match(im_left, im_right, left_key, right_key);
    match_size = (int)matches.size();
    total_key_num = (int)key_left_cube.size();
}
