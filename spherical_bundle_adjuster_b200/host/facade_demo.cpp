// facade_demo.cpp -- exercises the drop-in classes end to end on files written by the Python tests:
//   facade_demo <dir>   reads <dir>/{im.bin,desc1.bin,desc2.bin,key1.bin,key2.bin,meta.txt}
//                       writes <dir>/{strip.bin,face3.bin,matches.bin,rot.bin}
//                       and, when <dir>/sp_b1.bin is there (sp_b1, sp_b2: n x 3 doubles; sp_init: r0, t0, d0),
//                       runs solve_problem and writes <dir>/sp_out.bin (r, t, iterations of the 3 stages, d [n x 2])
// (tests/test_gpu_facade.py compares those with the oracle).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include "spherical_bundle_adjuster.hpp"

template <typename T> static std::vector<T> slurp(const std::string& p)
{
    std::ifstream f(p, std::ios::binary);
    std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::vector<T> v(raw.size() / sizeof(T));
    memcpy(v.data(), raw.data(), v.size() * sizeof(T));
    return v;
}
template <typename T> static void dump(const std::string& p, const T* d, size_t n)
{
    std::ofstream f(p, std::ios::binary);
    f.write((const char*)d, n * sizeof(T));
}

int main(int argc, char** argv)
{
    if (argc < 2) { std::cerr << "usage: facade_demo <dir>\n"; return 2; }
    const std::string dir = argv[1];
    int w, h, cs, n1, n2;
    { std::ifstream m(dir + "/meta.txt"); m >> w >> h >> cs >> n1 >> n2; }
    try {
        auto im = slurp<unsigned char>(dir + "/im.bin");
        cv::Mat erp(h, w, CV_8UC3, im.data());
        equi2cube e2c;
        cv::Mat strip = e2c.get_all(erp, cs);
        dump(dir + "/strip.bin", strip.data, strip.total() * 3);
        cv::Mat back = e2c.get_back(erp, cs);
        dump(dir + "/face3.bin", back.data, back.total() * 3);

        auto d1 = slurp<float>(dir + "/desc1.bin"), d2 = slurp<float>(dir + "/desc2.bin");
        auto k1 = slurp<float>(dir + "/key1.bin"), k2 = slurp<float>(dir + "/key2.bin");
        cv::Mat desc1(n1, 64, CV_32FC1, d1.data()), desc2(n2, 64, CV_32FC1, d2.data());
        std::vector<cv::KeyPoint> kp1(n1), kp2(n2);
        for (int i = 0; i < n1; i++) { kp1[i].pt.x = k1[2 * i]; kp1[i].pt.y = k1[2 * i + 1]; }
        for (int i = 0; i < n2; i++) { kp2[i].pt.x = k2[2 * i]; kp2[i].pt.y = k2[2 * i + 1]; }

        equi2cube_surf es;
        es.set_cube_size(cs);
        std::vector<cv::KeyPoint> left, right;
        std::vector<cv::DMatch> matches;
        es.match_and_lift(kp1, kp2, desc1, desc2, w, h, left, right, matches);
        std::vector<int> mm;
        for (auto& m : matches) { mm.push_back(m.queryIdx); mm.push_back(m.trainIdx); }
        dump(dir + "/matches.bin", mm.data(), mm.size());

        spherical_bundle_adjuster sba;
        double rot[3] = {0, 0, 0};
        sba_solve_summary s = sba.adjust_rotation(left, right, w, h, rot);
        double out[5] = {rot[0], rot[1], rot[2], (double)s.iterations, s.final_cost};
        dump(dir + "/rot.bin", out, 5);
        // spherical_surf: the four bands, one pitched crop, and the post-SURF part on supplied band keypoints/descriptors
        // (ss_k{l,r}.bin: 4 x m x 2 floats, ss_d{l,r}.bin: 4 x m x 64 floats)
        auto skl = slurp<float>(dir + "/ss_kl.bin");
        if (!skl.empty()) {
            spherical_surf ss;
            cv::Mat bands[4];
            ss.crop_bands(erp, bands);
            for (int b = 0; b < 4; b++) dump(dir + "/ss_band" + std::to_string(b) + ".bin", bands[b].data, bands[b].total() * 3);
            cv::Mat c45 = ss.crop_rotated_image(45, erp);
            dump(dir + "/ss_crop45.bin", c45.data, c45.total() * 3);
            cv::Mat rot = ss.eular2rot(cv::Vec3f(0, RAD(-45), 0));
            cv::Vec2i px = ss.rotate_pixel(cv::Vec2i(h / 2, w / 3), rot, w, h);
            auto skr = slurp<float>(dir + "/ss_kr.bin"), sdl = slurp<float>(dir + "/ss_dl.bin"), sdr = slurp<float>(dir + "/ss_dr.bin");
            const int m = (int)(skl.size() / 8);
            std::vector<cv::KeyPoint> kl[4], kr[4];
            cv::Mat dl[4], dr[4];
            for (int b = 0; b < 4; b++) {
                kl[b].resize(m); kr[b].resize(m);
                for (int i = 0; i < m; i++) {
                    kl[b][i].pt.x = skl[2 * (b * m + i)]; kl[b][i].pt.y = skl[2 * (b * m + i) + 1];
                    kr[b][i].pt.x = skr[2 * (b * m + i)]; kr[b][i].pt.y = skr[2 * (b * m + i) + 1];
                }
                dl[b] = cv::Mat(m, 64, CV_32FC1, sdl.data() + (size_t)b * m * 64);
                dr[b] = cv::Mat(m, 64, CV_32FC1, sdr.data() + (size_t)b * m * 64);
            }
            std::vector<cv::KeyPoint> L, R;
            std::vector<cv::DMatch> mm2;
            ss.lift_and_match(kl, kr, dl, dr, w, h, L, R, mm2);
            std::vector<float> o = {(float)px[0], (float)px[1]};
            for (size_t i = 0; i < mm2.size(); i++) {
                o.push_back((float)mm2[i].queryIdx); o.push_back((float)mm2[i].trainIdx);
                o.push_back(L[i].pt.x); o.push_back(L[i].pt.y); o.push_back(R[i].pt.x); o.push_back(R[i].pt.y);
            }
            dump(dir + "/ss_out.bin", o.data(), o.size());
        }
        auto sb1 = slurp<double>(dir + "/sp_b1.bin");
        if (!sb1.empty()) {
            auto sb2 = slurp<double>(dir + "/sp_b2.bin"), init = slurp<double>(dir + "/sp_init.bin");
            const int n = (int)(sb1.size() / 3);
            std::vector<cv::Point3d> L(n), R(n);
            for (int i = 0; i < n; i++) { L[i] = cv::Point3d(sb1[3 * i], sb1[3 * i + 1], sb1[3 * i + 2]); R[i] = cv::Point3d(sb2[3 * i], sb2[3 * i + 1], sb2[3 * i + 2]); }
            double r0[3] = {init[0], init[1], init[2]}, t0[3] = {init[3], init[4], init[5]};
            std::vector<std::array<double, 2>> init_d(n);
            for (auto& d : init_d) d[0] = d[1] = init[6];
            cv::Vec3f Rg, Tg;
            std::srand(1);                                    // the reference never seeds: std::rand starts from 1
            sba.initial_guess(w, h, L, R, Rg, Tg, n);
            float ig[6] = {Rg[0], Rg[1], Rg[2], Tg[0], Tg[1], Tg[2]};
            dump(dir + "/ig_out.bin", ig, 6);
            sba_solver_options opt;
            sba.solve_problem(opt, L, R, r0, t0, init_d, n);
            std::vector<double> o = {r0[0], r0[1], r0[2], t0[0], t0[1], t0[2], (double)sba.stage_summaries[0].iterations,
                                     (double)sba.stage_summaries[1].iterations, (double)sba.stage_summaries[2].iterations};
            for (auto& d : init_d) { o.push_back(d[0]); o.push_back(d[1]); }
            dump(dir + "/sp_out.bin", o.data(), o.size());
        }
        std::printf("facade_demo: %zu matches, rotation %.9f %.9f %.9f, %d LM iterations\n", matches.size(), rot[0], rot[1], rot[2], s.iterations);
    } catch (const std::exception& ex) {
        std::cerr << "facade_demo failed: " << ex.what() << "\n";
        return 1;
    }
    return 0;
}
