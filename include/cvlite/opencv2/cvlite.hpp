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
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

typedef int64_t int64;
typedef unsigned char uchar;

#define CV_CN_SHIFT 3
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) (((depth)&7) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_32SC2 CV_MAKETYPE(CV_32S, 2)

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
template <typename T, int N> inline Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b)
{
    Vec<T, N> r;
    for (int i = 0; i < N; i++) r.val[i] = a.val[i] - b.val[i];
    return r;
}

template <typename T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
};
typedef Rect_<int> Rect;

// element type -> type code, for the typed matrices below
template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8UC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };
template <> struct DataType<Vec<int, 2>> { enum { type = CV_32SC2 }; };
template <> struct DataType<Vec<uchar, 3>> { enum { type = CV_8UC3 }; };

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
    template <typename T> T& at(int r, int c) { return ((T*)data)[(size_t)r * cols + c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)data)[(size_t)r * cols + c]; }
    // Sub-matrix header sharing the parent's storage.  Only full-width row bands are representable in this
    // continuous-only carrier (that is the one shape the reference takes: spherical_surf.cpp:132,139).
    Mat operator()(const Rect& roi) const
    {
        if (roi.x != 0 || roi.width != cols || roi.y < 0 || roi.y + roi.height > rows) { std::cerr << "cvlite: only full-width row bands\n"; std::abort(); }
        Mat m;
        m.rows = roi.height; m.cols = cols; m.type_ = type_; m.store_ = store_;
        m.data = data + (size_t)roi.y * cols * elemSize();
        return m;
    }

private:
    int type_;
    std::shared_ptr<std::vector<uchar>> store_;
};

// Typed matrix with the `Mat_<T>(r, c) << a, b, c ...` initialiser (opencv2/core/mat.hpp).
template <typename T> class Mat_;
template <typename T> class MatCommaInitializer_ {
public:
    MatCommaInitializer_(Mat_<T>* m) : m_(m), k_(0) {}
    template <typename T2> MatCommaInitializer_<T>& operator,(T2 v)
    {
        ((T*)m_->data)[k_++] = (T)v;
        return *this;
    }
    operator Mat_<T>() const { return *m_; }
    operator Mat() const { return *m_; }
    Mat_<T>* m_;
    size_t k_;
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
};
template <typename T, typename T2> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, T2 v)
{
    MatCommaInitializer_<T> ci(const_cast<Mat_<T>*>(&m));
    return (ci, v);
}
typedef Mat_<Vec2i> Mat2i;
typedef Mat_<double> Mat1d;

// Matrix product of CV_64FC1 matrices (the reference multiplies 3x3 rotation factors, spherical_surf.cpp:42).
inline Mat operator*(const Mat& a, const Mat& b)
{
    if (a.type() != CV_64FC1 || b.type() != CV_64FC1 || a.cols != b.rows) { std::cerr << "cvlite: operator* is CV_64FC1 only\n"; std::abort(); }
    Mat c(a.rows, b.cols, CV_64FC1);
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.at<double>(i, k) * b.at<double>(k, j);
            c.at<double>(i, j) = s;
        }
    return c;
}

// Vertical concatenation of equally wide, same-type matrices (cv::vconcat, array form).
inline void vconcat(const Mat* src, size_t n, Mat& dst)
{
    int rows = 0, cols = 0, type = 0;
    for (size_t i = 0; i < n; i++) if (!src[i].empty()) { rows += src[i].rows; cols = src[i].cols; type = src[i].type(); }
    Mat out(rows, cols, type);
    size_t off = 0;
    for (size_t i = 0; i < n; i++) {
        if (src[i].empty()) continue;
        const size_t bytes = src[i].total() * src[i].elemSize();
        std::memcpy(out.data + off, src[i].data, bytes);
        off += bytes;
    }
    dst = out;
}

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
