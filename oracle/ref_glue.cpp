// ref_glue.cpp -- C entry points around the REFERENCE's own equi2cube / equi2cube_surf classes.
// TEST INFRASTRUCTURE ONLY (same rules as sba_oracle.c).
//
// oracle/Makefile compiles /root/reference/equi2cube.cpp and /root/reference/equi2cube_surf.cpp
// from where they lie (never copied) against the cv type shim in include/cvlite, and links this
// file in to expose them through a C ABI that tests and bench.py can load with ctypes.
// feature_matcher's methods are referenced by equi2cube_surf::do_all but need SURF (non-free,
// absent); they are defined here as aborting stubs so the library links -- do_all is never called.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "equi2cube.hpp"
#include "equi2cube_surf.hpp"
// crop_rotated_image and rotate_keypoint are private members of the reference class (spherical_surf.hpp:19-20);
// the checker reaches them without touching the reference source
#define private public
#include "spherical_surf.hpp"
#undef private

static void unavailable(const char* what)
{
    std::fprintf(stderr, "oracle/_ref: %s needs OpenCV xfeatures2d (SURF), which is not available\n", what);
    std::abort();
}
void feature_matcher::init() {}
void feature_matcher::deinit() {}
std::vector<cv::KeyPoint> feature_matcher::detect_key_point(const cv::Mat&) { unavailable("detect_key_point"); return {}; }
cv::Mat feature_matcher::comput_descriptor(const cv::Mat&, std::vector<cv::KeyPoint>&) { unavailable("comput_descriptor"); return {}; }
std::vector<cv::DMatch> feature_matcher::match_two_image(const cv::Mat&, const cv::Mat&) { unavailable("match_two_image"); return {}; }
cv::Mat feature_matcher::draw_match(const cv::Mat&, const cv::Mat&, const std::vector<cv::KeyPoint>&, const std::vector<cv::KeyPoint>&) { unavailable("draw_match"); return {}; }

extern "C" {

// The reference's get_bottom reads im_data[h*w + col] for the exact bottom-centre pixel when the
// polar angle rounds to pi (equi2cube.cpp:268-275).  Callers pass an image buffer with one extra
// padding row (h+1 rows allocated, `h` declared) so that read stays inside the allocation; the
// padding row replicates row h-1, which is the clamp the product documents.
void ref_equi2cube_all(const unsigned char* im_padded, int w, int h, int cs, int nthreads, unsigned char* strip)
{
    cv::Mat im(h, w, CV_8UC3, (void*)im_padded);
    equi2cube e;
    omp_set_num_threads(nthreads);  // what equi2cube::set_omp does (equi2cube.cpp:8) minus its stdout print
    cv::Mat out = e.get_all(im, cs);
    std::memcpy(strip, out.data, (size_t)cs * 6 * cs * 3);
}

// face: 0 left, 1 front, 2 right, 3 back, 4 top, 5 bottom (strip order of get_all)
void ref_equi2cube_face(const unsigned char* im_padded, int w, int h, int cs, int face, unsigned char* out)
{
    cv::Mat im(h, w, CV_8UC3, (void*)im_padded);
    equi2cube e;
    cv::Mat f;
    switch (face) {
    case 0: f = e.get_left(im, cs); break;
    case 1: f = e.get_front(im, cs); break;
    case 2: f = e.get_right(im, cs); break;
    case 3: f = e.get_back(im, cs); break;
    case 4: f = e.get_top(im, cs); break;
    default: f = e.get_bottom(im, cs); break;
    }
    std::memcpy(out, f.data, (size_t)cs * cs * 3);
}

void ref_cube2equi_points(const float* xy_in, int n, int cs, int w, int h, float* xy_out)
{
    // equi2cube_surf has no user constructor and its ctor-less members stay unset; cube2equi_pixel
    // takes everything it needs as arguments (equi2cube_surf.cpp:19).
    equi2cube_surf s;
    for (int k = 0; k < n; k++) {
        cv::Point2f in(xy_in[2 * k], xy_in[2 * k + 1]), out;
        s.cube2equi_pixel(in, out, cs, w, h);
        xy_out[2 * k] = out.x;
        xy_out[2 * k + 1] = out.y;
    }
}

// ---- spherical_surf (spherical_surf.cpp:17-123) ---------------------------------------------------

// theta: Euler angles as the float triple the reference passes; R: 3x3 row-major doubles.
void ref_eular2rot(const float theta[3], double R[9])
{
    spherical_surf s;
    cv::Mat m = s.eular2rot(cv::Vec3f(theta[0], theta[1], theta[2]));
    std::memcpy(R, m.data, 9 * sizeof(double));
}

// rc_in / rc_out: n (row, col) integer pairs; the rotation is eular2rot(0, RAD(pitch_deg), 0) like both callers build it.
void ref_rotate_pixels(const int* rc_in, int n, float pitch_deg, int w, int h, int* rc_out)
{
    spherical_surf s;
    cv::Mat rot = s.eular2rot(cv::Vec3f(0, RAD(pitch_deg), 0));
    for (int k = 0; k < n; k++) {
        cv::Vec2i o = s.rotate_pixel(cv::Vec2i(rc_in[2 * k], rc_in[2 * k + 1]), rot, w, h);
        rc_out[2 * k] = o[0];
        rc_out[2 * k + 1] = o[1];
    }
}

// out: (h/4) x w x 3.  Pixels whose source falls outside the image are left as the reference leaves them
// (unwritten); the shim's Mat storage starts zeroed, so they read 0 here.
void ref_crop_rotated_image(const unsigned char* im, int w, int h, float pitch_deg, int nthreads, unsigned char* out)
{
    cv::Mat m(h, w, CV_8UC3, (void*)im);
    spherical_surf s;
    omp_set_num_threads(nthreads);
    cv::Mat o = s.crop_rotated_image(pitch_deg, m);
    std::memcpy(out, o.data, (size_t)(h / 4) * w * 3);
}

// xy: n keypoints (x, y) in the cropped band's coordinates, rotated in place (spherical_surf.cpp:111-123).
void ref_rotate_keypoints(float pitch_inv_deg, float* xy, int n, int w, int h)
{
    std::vector<cv::KeyPoint> key(n);
    for (int k = 0; k < n; k++) { key[k].pt.x = xy[2 * k]; key[k].pt.y = xy[2 * k + 1]; }
    spherical_surf s;
    s.rotate_keypoint(pitch_inv_deg, key, w, h);
    for (int k = 0; k < n; k++) { xy[2 * k] = key[k].pt.x; xy[2 * k + 1] = key[k].pt.y; }
}

}  // extern "C"
