// Facade: pitched-band front-end.  Replaces spherical_surf.cpp:8-232 of the reference.
#include "spherical_surf.hpp"

#include <cmath>

#include "sba_host_ctx.hpp"

static const float kPitch[4] = {45.f, 0.f, -45.f, -90.f};   // spherical_surf.cpp:137-143; band 1 is the plain crop im(roi)

void spherical_surf::set_omp(int num_proc) { this->num_proc = num_proc; }

cv::Mat spherical_surf::eular2rot(cv::Vec3f theta)
{
    cv::Mat R(3, 3, CV_64FC1);
    const float th[3] = {theta[0], theta[1], theta[2]};
    sba_host::check(sba_eular2rot(th, (double*)R.data));
    return R;
}

cv::Vec2i spherical_surf::rotate_pixel(const cv::Vec2i& in_vec, cv::Mat& rot_mat, int width, int height)
{
    const int32_t in[2] = {in_vec[0], in_vec[1]};
    int32_t out[2];
    sba_host::check(sba_rotate_pixels_mat(sba_host::ctx(), in, 1, (const double*)rot_mat.data, width, height, out, SBA_MEM_HOST));
    return cv::Vec2i(out[0], out[1]);
}

cv::Mat spherical_surf::crop_rotated_image(float pitch_rot, const cv::Mat& im)
{
    cv::Mat out(im.rows / 4, im.cols, im.type());
    sba_host::check(sba_crop_rotated_image(sba_host::ctx(), im.data, im.cols, im.rows, 1, pitch_rot, out.data, SBA_MEM_HOST));
    return out;
}

void spherical_surf::crop_bands(const cv::Mat& im, cv::Mat bands[4])
{
    const int bh = im.rows / 4;
    cv::Mat all(4 * bh, im.cols, im.type());
    sba_host::check(sba_spherical_crops(sba_host::ctx(), im.data, im.cols, im.rows, 1, all.data, SBA_MEM_HOST));
    for (int b = 0; b < 4; b++) {
        bands[b] = cv::Mat(bh, im.cols, im.type());
        std::memcpy(bands[b].data, all.data + (size_t)b * bh * im.cols * im.elemSize(), (size_t)bh * im.cols * im.elemSize());
    }
}

void spherical_surf::rotate_keypoint(float pitch_rot_inv, std::vector<cv::KeyPoint>& key, int width, int height)
{
    const int n = (int)key.size();
    if (n == 0) return;
    std::vector<float> xy(2 * (size_t)n);
    for (int i = 0; i < n; i++) { xy[2 * i] = key[i].pt.x; xy[2 * i + 1] = key[i].pt.y; }
    sba_host::check(sba_rotate_keypoints(sba_host::ctx(), xy.data(), n, pitch_rot_inv, width, height, SBA_MEM_HOST));
    for (int i = 0; i < n; i++) { key[i].pt.x = xy[2 * i]; key[i].pt.y = xy[2 * i + 1]; }
}

void spherical_surf::lift_and_match(std::vector<cv::KeyPoint> key_left[4], std::vector<cv::KeyPoint> key_right[4], const cv::Mat desc_left[4],
                                    const cv::Mat desc_right[4], int im_width, int im_height, std::vector<cv::KeyPoint>& left_key,
                                    std::vector<cv::KeyPoint>& right_key, std::vector<cv::DMatch>& matches)
{
    // spherical_surf.cpp:180-193: band keypoints back to ERP pixels (bands 0, 2, 3 through the crop's mapping,
    // band 1 by the row offset of the plain crop)
    std::vector<cv::KeyPoint>* sides[2] = {key_left, key_right};
    for (auto* key : sides)
        for (int b = 0; b < 4; b++) {
            if (b == 1) for (auto& k : key[1]) k.pt.y = k.pt.y + im_height * 3 / 8;
            else rotate_keypoint(kPitch[b], key[b], im_width, im_height);
        }
    // :195-211: concatenate in band order.  (The reference appends to members it never clears, :196-204, so its
    // object cannot be reused; locals here.)
    std::vector<cv::KeyPoint> all_left, all_right;
    for (int b = 0; b < 4; b++) {
        all_left.insert(all_left.end(), key_left[b].begin(), key_left[b].end());
        all_right.insert(all_right.end(), key_right[b].begin(), key_right[b].end());
    }
    cv::Mat d_left, d_right;
    cv::vconcat(desc_left, 4, d_left);
    cv::vconcat(desc_right, 4, d_right);
    feature_matcher fm;
    matches = fm.match_two_image(d_left, d_right);          // :214
    left_key.resize(matches.size());
    right_key.resize(matches.size());
    for (size_t i = 0; i < matches.size(); i++) {            // :216-222
        left_key[i] = all_left[matches[i].queryIdx];
        right_key[i] = all_right[matches[i].trainIdx];
    }
}

void spherical_surf::do_all(const cv::Mat& im_left, const cv::Mat& im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num)
{
    // spherical_surf.cpp:125-232
    const int im_width = im_left.cols, im_height = im_left.rows;
    cv::Mat bands_left[4], bands_right[4];
    crop_bands(im_left, bands_left);
    crop_bands(im_right, bands_right);

    feature_matcher fm;
    std::vector<cv::KeyPoint> key_left[4], key_right[4];
    cv::Mat desc_left[4], desc_right[4];
    for (int b = 0; b < 4; b++) key_left[b] = fm.detect_key_point(bands_left[b]);
    for (int b = 0; b < 4; b++) key_right[b] = fm.detect_key_point(bands_right[b]);
    for (int b = 0; b < 4; b++) desc_left[b] = fm.comput_descriptor(bands_left[b], key_left[b]);
    for (int b = 0; b < 4; b++) desc_right[b] = fm.comput_descriptor(bands_right[b], key_right[b]);
    total_key_num = 0;
    for (int b = 0; b < 4; b++) total_key_num += (int)key_left[b].size();

    std::vector<cv::DMatch> matches;
    lift_and_match(key_left, key_right, desc_left, desc_right, im_width, im_height, left_key, right_key, matches);
    match_output = fm.draw_match(im_left, im_right, left_key, right_key);
    match_size = (int)matches.size();
}
