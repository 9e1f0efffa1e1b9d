// Facade over sba_equi2cube*.  Replaces equi2cube.cpp:12-302 of the reference.
#include "equi2cube.hpp"

#include "sba_host_ctx.hpp"

// The reference calls the process-global omp_set_num_threads here (equi2cube.cpp:8); there is nothing
// to configure on the GPU path, and leaving the caller's OpenMP state alone is the safer behaviour.
void equi2cube::set_omp(int) {}

cv::Mat equi2cube::face(const cv::Mat& im, int cube_size, int face_id)
{
    cv::Mat out(cube_size, cube_size, CV_8UC3);
    sba_host::check(sba_equi2cube_face(sba_host::ctx(), im.data, im.cols, im.rows, cube_size, face_id, out.data, SBA_MEM_HOST));
    return out;
}

// face ids follow get_all's strip order (equi2cube.cpp:293-298): left, front, right, back, top, bottom
cv::Mat equi2cube::get_left(const cv::Mat& im, int cube_size) { return face(im, cube_size, 0); }
cv::Mat equi2cube::get_front(const cv::Mat& im, int cube_size) { return face(im, cube_size, 1); }
cv::Mat equi2cube::get_right(const cv::Mat& im, int cube_size) { return face(im, cube_size, 2); }
cv::Mat equi2cube::get_back(const cv::Mat& im, int cube_size) { return face(im, cube_size, 3); }
cv::Mat equi2cube::get_top(const cv::Mat& im, int cube_size) { return face(im, cube_size, 4); }
cv::Mat equi2cube::get_bottom(const cv::Mat& im, int cube_size) { return face(im, cube_size, 5); }

cv::Mat equi2cube::get_all(const cv::Mat& im, int cube_size)
{
    cv::Mat out(cube_size, 6 * cube_size, CV_8UC3);
    sba_host::check(sba_equi2cube(sba_host::ctx(), im.data, im.cols, im.rows, 1, cube_size, out.data, SBA_MEM_HOST));
    return out;
}
