// Facade over sba_knn2_ratio.  Replaces feature_matcher.cpp:42-59 (the matcher); detect / describe /
// draw (feature_matcher.cpp:26-40, :61-86) remain OpenCV host code when OpenCV is present.
#include "feature_matcher.hpp"

#include <stdexcept>

#include "sba_host_ctx.hpp"

void feature_matcher::init()
{
#if defined(SBA_HAVE_OPENCV_XFEATURES2D)
    detector = cv::xfeatures2d::SURF::create();              // feature_matcher.cpp:13
    descriptor_extractor = cv::xfeatures2d::SURF::create();  // feature_matcher.cpp:15
#endif
    // the reference creates a FLANN matcher here (feature_matcher.cpp:16); matching is the GPU's job now
}

void feature_matcher::deinit() {}

std::vector<cv::KeyPoint> feature_matcher::detect_key_point(const cv::Mat& image)
{
#if defined(SBA_HAVE_OPENCV_XFEATURES2D)
    std::vector<cv::KeyPoint> key_point;
    detector->detect(image, key_point);
    return key_point;
#else
    (void)image;
    throw std::runtime_error("feature_matcher::detect_key_point needs OpenCV xfeatures2d (SURF); build with real OpenCV");
#endif
}

cv::Mat feature_matcher::comput_descriptor(const cv::Mat& image, std::vector<cv::KeyPoint>& key_point)
{
#if defined(SBA_HAVE_OPENCV_XFEATURES2D)
    cv::Mat descriptors;
    descriptor_extractor->compute(image, key_point, descriptors);
    return descriptors;
#else
    (void)image; (void)key_point;
    throw std::runtime_error("feature_matcher::comput_descriptor needs OpenCV xfeatures2d (SURF); build with real OpenCV");
#endif
}

// kNN(k=2) + Lowe ratio 0.3, survivors in query order (feature_matcher.cpp:42-59).
std::vector<cv::DMatch> feature_matcher::match_two_image(const cv::Mat& descriptor1, const cv::Mat& descriptor2)
{
    const float ratio_thresh = 0.3f;
    const int nq = descriptor1.rows, nt = descriptor2.rows;
    const int dim = nq ? descriptor1.cols : descriptor2.cols;
    std::vector<cv::DMatch> good_matches;
    if (nq == 0) return good_matches;
    // the library reads the descriptors as dense fp32 rows: say so instead of reinterpreting whatever arrives
    // (the reference hands cv::Mat to OpenCV's matcher, which converts / rejects on its own)
    if (descriptor1.type() != CV_32FC1 || (nt && descriptor2.type() != CV_32FC1) || !descriptor1.isContinuous() || !descriptor2.isContinuous() ||
        (nt && descriptor2.cols != dim) || (dim != 64 && dim != 128))
        throw std::runtime_error("feature_matcher::match_two_image: descriptors must be continuous CV_32F matrices with 64 or 128 columns (SURF)");
    std::vector<int32_t> qi(nq), ti(nq);
    std::vector<float> dist(nq);
    int32_t n = 0;
    sba_host::check(sba_knn2_ratio(sba_host::ctx(), (const float*)descriptor1.data, nq, (const float*)descriptor2.data, nt, dim, ratio_thresh,
                                   qi.data(), ti.data(), dist.data(), &n, nullptr, nullptr, SBA_MEM_HOST, SBA_MATCH_AUTO));
    good_matches.reserve(n);
    for (int i = 0; i < n; i++) good_matches.push_back(cv::DMatch(qi[i], ti[i], 0, dist[i]));   // imgIdx 0 like BFMatcher
    return good_matches;
}

cv::Mat feature_matcher::draw_match(const cv::Mat& im_left, const cv::Mat& im_right, const std::vector<cv::KeyPoint>& key_left, const std::vector<cv::KeyPoint>& key_right)
{
#if defined(SBA_HAVE_OPENCV_XFEATURES2D)
    // feature_matcher.cpp:61-86, unchanged (debug visualisation, host OpenCV)
    cv::Mat im_left_gray, im_right_gray;
    cv::cvtColor(im_left, im_left_gray, cv::COLOR_RGB2GRAY);
    cv::cvtColor(im_right, im_right_gray, cv::COLOR_RGB2GRAY);
    cv::Mat im_overlap(im_left.rows, im_left.cols, im_left.type());
    cv::Mat chan[3] = {im_left_gray, im_right_gray, cv::Mat::zeros(im_left.rows, im_left.cols, CV_8UC1)};
    cv::merge(chan, 3, im_overlap);
    int match_size = (int)key_left.size();
    for (int i = 0; i < match_size; i++) {
        cv::Mat rgb, hsv(1, 1, CV_8UC3, cv::Scalar(i * (180.0 / match_size), 180, 150));
        cv::cvtColor(hsv, rgb, cv::COLOR_HSV2BGR);
        cv::line(im_overlap, key_left[i].pt, key_right[i].pt, cv::Scalar(rgb.data[0], rgb.data[1], rgb.data[2]), 5);
    }
    return im_overlap;
#else
    (void)im_right; (void)key_left; (void)key_right;
    return cv::Mat(im_left.rows, im_left.cols, im_left.type());   // no drawing primitives without OpenCV imgproc
#endif
}

void feature_matcher::do_all(const cv::Mat& im_left, const cv::Mat& im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num)
{
    // feature_matcher.cpp:88-128
    key_point_left = detect_key_point(im_left);
    key_point_right = detect_key_point(im_right);
    descriptor_left = comput_descriptor(im_left, key_point_left);
    descriptor_right = comput_descriptor(im_right, key_point_right);
    matches = match_two_image(descriptor_left, descriptor_right);
    std::vector<cv::KeyPoint> valid_key_left(matches.size()), valid_key_right(matches.size());
    for (size_t i = 0; i < matches.size(); i++) {
        valid_key_left[i] = key_point_left[matches[i].queryIdx];
        valid_key_right[i] = key_point_right[matches[i].trainIdx];
    }
    match_output = draw_match(im_left, im_right, valid_key_left, valid_key_right);
    left_key = valid_key_left;
    right_key = valid_key_right;
    match_size = (int)matches.size();
    total_key_num = (int)key_point_left.size();
}
