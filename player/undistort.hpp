// Image undistortion of the dataset readers, on the CPU like in the reference player:
//   EuRoC : cv::undistort(img, K, D) with the radtan model, K and D handed over as float32
//           (xrslam-pc/player/src/IO/euroc_dataset_reader.cpp:62-72)
//   TUM-VI: xrslam::extra::ImageUndistorter(..., "equidistant") + cv::remap(INTER_LINEAR, BORDER_CONSTANT)
//           (xrslam-extra/include/xrslam/extra/image_undistorter.h:13-97, tum_dataset_reader.cpp:62-76)
// Both end in OpenCV's fixed-point bilinear remap: source coordinates rounded to 1/32 px, weights
// scaled by 2^15, result (sum + 2^14) >> 15, zero outside the image.  That arithmetic is restated
// here so that the player needs no OpenCV.  (The same remap as a CUDA kernel inside the library is
// SURVEY.md row N2's remaining item.)
#pragma once
#include "png_reader.hpp"
#include <cmath>
#include <string>

namespace xplayer {

struct Undistorter {
    int width = 0, height = 0;
    std::vector<short> ix, iy;          // integer source coordinates
    std::vector<unsigned short> frac;   // (fy << 5) | fx, 5-bit fractions

    static int round_half_even(double v) { return (int)std::nearbyint(v); }     // cvRound / saturate_cast<int>(double)
    static short sat16(int v) { return (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v)); }

    // K = fx fy cx cy; D = k1 k2 p1 p2 (radtan) or k1 k2 k3 k4 (equidistant)
    void build(int w, int h, const double K[4], const double D[4], const std::string &model, bool float_params) {
        width = w; height = h;
        ix.assign((size_t)w * h, 0); iy.assign((size_t)w * h, 0); frac.assign((size_t)w * h, 0);
        double k[4] = {K[0], K[1], K[2], K[3]}, d[4] = {D[0], D[1], D[2], D[3]};
        if (float_params) { for (int i = 0; i < 4; ++i) { k[i] = (double)(float)k[i]; d[i] = (double)(float)d[i]; } }
        const double fx = k[0], fy = k[1], cx = k[2], cy = k[3];
        for (int v = 0; v < h; ++v)
            for (int u = 0; u < w; ++u) {
                double sx, sy;
                if (model == "radtan") {
                    // cv::initUndistortRectifyMap with R = I and the same camera matrix
                    const double x = (u - cx) / fx, y = (v - cy) / fy;
                    const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
                    const double kr = 1 + (d[1] * r2 + d[0]) * r2;
                    const double xd = x * kr + d[2] * _2xy + d[3] * (r2 + 2 * x2);
                    const double yd = y * kr + d[2] * (r2 + 2 * y2) + d[3] * _2xy;
                    sx = fx * xd + cx; sy = fy * yd + cy;
                } else {
                    // ImageUndistorter::distort_pixel, "equidistant" (image_undistorter.h:66-84); the map is stored as float32
                    const double x = (u - cx) / fx, y = (v - cy) / fy;
                    const double r = std::sqrt(x * x + y * y);
                    if (r < 1e-10) { sx = u; sy = v; }
                    else {
                        const double th = std::atan(r), t2 = th * th, t4 = t2 * t2, t6 = t2 * t4, t8 = t4 * t4;
                        const double thd = th * (1 + d[0] * t2 + d[1] * t4 + d[2] * t6 + d[3] * t8);
                        const double sc = (r > 1e-8) ? thd / r : 1.0;
                        sx = fx * (x * sc) + cx; sy = fy * (y * sc) + cy;
                    }
                    sx = (double)(float)sx; sy = (double)(float)sy;
                }
                const int qx = round_half_even(sx * 32.0), qy = round_half_even(sy * 32.0);
                const size_t i = (size_t)v * w + u;
                ix[i] = sat16(qx >> 5); iy[i] = sat16(qy >> 5);
                frac[i] = (unsigned short)(((qy & 31) << 5) | (qx & 31));
            }
    }

    void apply(const GrayImage &src, GrayImage &dst) const {
        dst.width = width; dst.height = height; dst.data.assign((size_t)width * height, 0);
        const int sw = src.width, sh = src.height;
        auto at = [&](int x, int y) -> int { return (x >= 0 && x < sw && y >= 0 && y < sh) ? src.data[(size_t)y * sw + x] : 0; };
        for (size_t i = 0; i < (size_t)width * height; ++i) {
            const int x = ix[i], y = iy[i], fx = frac[i] & 31, fy = frac[i] >> 5;
            const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
            const int s = at(x, y) * w00 + at(x + 1, y) * w01 + at(x, y + 1) * w10 + at(x + 1, y + 1) * w11;
            dst.data[i] = (uint8_t)((s + (1 << 14)) >> 15);
        }
    }
};

}  // namespace xplayer
