// Drop-in for the reference's feature_matcher.hpp:24-49.  match_two_image runs on the GPU
// (sba_knn2_ratio: exact L2 kNN k=2, BFMatcher semantics, ratio 0.3).  SURF detect/describe stay with
// OpenCV xfeatures2d on the host: with real OpenCV on the include path (SBA_HAVE_OPENCV_XFEATURES2D)
// they are the reference's own two lines; without it (the cvlite type shim) they throw.
#pragma once
#include <vector>

#include "opencv2/core.hpp"
#if defined(SBA_HAVE_OPENCV_XFEATURES2D)
#include "opencv2/features2d.hpp"
#include "opencv2/imgproc.hpp"
#include "opencv2/xfeatures2d.hpp"
#endif

class feature_matcher
{
    public:
    void init();
    void deinit();
    feature_matcher(){ init(); }
    ~feature_matcher() { deinit(); }

    std::vector<cv::KeyPoint> detect_key_point(const cv::Mat &image);
    cv::Mat comput_descriptor(const cv::Mat &image, std::vector<cv::KeyPoint> &key_point);
    std::vector<cv::DMatch> match_two_image(const cv::Mat &descriptor1, const cv::Mat &descriptor2);
    cv::Mat draw_match(const cv::Mat& im_left, const cv::Mat& im_right, const std::vector<cv::KeyPoint>& key_left, const std::vector<cv::KeyPoint>& key_right);

    void do_all(const cv::Mat &im_left, const cv::Mat &im_right, std::vector<cv::KeyPoint>& left_key, std::vector<cv::KeyPoint>& right_key, int& match_size, cv::Mat& match_output, int& total_key_num);

    private:
    cv::Ptr<cv::Feature2D> detector;
    cv::Ptr<cv::Feature2D> descriptor_extractor;

    std::vector<cv::KeyPoint> key_point_left;
    std::vector<cv::KeyPoint> key_point_right;
    cv::Mat descriptor_left;
    cv::Mat descriptor_right;
    std::vector<cv::DMatch> matches;
};
