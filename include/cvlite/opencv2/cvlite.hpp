// cvlite.hpp -- minimal stand-in for the handful of OpenCV types the reference's public
// interfaces mention (cv::Mat, cv::KeyPoint, cv::DMatch, cv::Point*, cv::Vec*).
//
// OpenCV's C++ headers are not installed in the build image, so the drop-in facade classes
// (spherical_bundle_adjuster_b200/host/) and the oracle/_ref build of the reference's own
// equi2cube.cpp compile against this shim.  With real OpenCV present, put its include dir
// BEFORE include/cvlite on the include path and this file is never seen.
//
// Only data carriers live here -- no algorithm.  Field names and layouts follow
// opencv2/core/types.hpp and opencv2/core/mat.hpp so code written against cv:: compiles unchanged.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

typedef int64_t int64;
typedef unsigned char uchar;

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) (((depth)&7) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

typedef std::string String;

template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(); }
    Vec(T a, T b) { static_assert(N == 2, "Vec2 ctor"); val[0] = a; val[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, "Vec3 ctor"); val[0] = a; val[1] = b; val[2] = c; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
typedef Vec<uchar, 3> Vec3b;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 2> Vec2d;
typedef Vec<float, 3> Vec3f;
typedef Vec<int, 2> Vec2i;

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point_<int> Point;

template <typename T> struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<double> Point3d;
typedef Point3_<float> Point3f;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

struct DMatch {
    int queryIdx, trainIdx, imgIdx;
    float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
    DMatch(int q, int t, int im, float d) : queryIdx(q), trainIdx(t), imgIdx(im), distance(d) {}
    bool operator<(const DMatch& m) const { return distance < m.distance; }
};

// Dense, continuous, row-major matrix with shared ownership -- the subset of cv::Mat the
// reference's interfaces rely on (rows, cols, data, type(), elemSize(), clone(), empty()).
class Mat {
public:
    int rows, cols;
    uchar* data;
    Mat() : rows(0), cols(0), data(nullptr), type_(0) {}
    Mat(int rows_, int cols_, int type) { create(rows_, cols_, type); }
    // non-owning view over caller memory (cv::Mat(rows, cols, type, void*) semantics)
    Mat(int rows_, int cols_, int type, void* ext) : rows(rows_), cols(cols_), data((uchar*)ext), type_(type) {}
    void create(int rows_, int cols_, int type)
    {
        rows = rows_; cols = cols_; type_ = type;
        store_ = std::make_shared<std::vector<uchar>>((size_t)rows * cols * elemSize());
        data = store_->data();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    int depth() const { return type_ & 7; }
    size_t elemSize() const
    {
        static const int sz[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return (size_t)sz[depth()] * channels();
    }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || total() == 0; }
    bool isContinuous() const { return true; }
    Mat clone() const
    {
        Mat m(rows, cols, type_);
        if (!empty()) std::memcpy(m.data, data, total() * elemSize());
        return m;
    }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * cols * elemSize()); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * cols * elemSize()); }

private:
    int type_;
    std::shared_ptr<std::vector<uchar>> store_;
};

// Horizontal concatenation of equally tall, same-type matrices (cv::hconcat).
inline void hconcat(const std::vector<Mat>& src, Mat& dst)
{
    if (src.empty()) { dst = Mat(); return; }
    int rows = src[0].rows, cols = 0;
    for (const Mat& m : src) cols += m.cols;
    Mat out(rows, cols, src[0].type());
    size_t es = src[0].elemSize();
    for (int r = 0; r < rows; r++) {
        size_t off = 0;
        for (const Mat& m : src) {
            std::memcpy(out.data + ((size_t)r * cols) * es + off, m.data + (size_t)r * m.cols * es, (size_t)m.cols * es);
            off += (size_t)m.cols * es;
        }
    }
    dst = out;
}

inline int64 getTickCount()
{
    return (int64)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline double getTickFrequency() { return 1e9; }

template <typename T> using Ptr = std::shared_ptr<T>;

// Opaque algorithm handles: named only so that headers declaring cv::Ptr<cv::Feature2D> members
// (feature_matcher.hpp:41-43) parse.  SURF itself stays with real OpenCV (out of scope).
class Feature2D {};
class DescriptorMatcher {};

}  // namespace cv
