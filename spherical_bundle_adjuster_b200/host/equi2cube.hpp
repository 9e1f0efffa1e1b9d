// Drop-in for the reference's equi2cube.hpp:20-32 -- same class, same signatures; the remap runs on the
// GPU through sba_equi2cube / sba_equi2cube_face (include/sba_b200.h).
#pragma once
#include "opencv2/core.hpp"

class equi2cube
{
    public:
    void set_omp(int num_proc);   // kept for source compatibility; the GPU path has no OpenMP threads
    cv::Mat get_back(const cv::Mat& im, int cube_size);
    cv::Mat get_front(const cv::Mat& im, int cube_size);
    cv::Mat get_left(const cv::Mat& im, int cube_size);
    cv::Mat get_right(const cv::Mat& im, int cube_size);
    cv::Mat get_top(const cv::Mat& im, int cube_size);
    cv::Mat get_bottom(const cv::Mat& im, int cube_size);
    cv::Mat get_all(const cv::Mat& im, int cube_size);
    private:
    cv::Mat face(const cv::Mat& im, int cube_size, int face_id);
};
