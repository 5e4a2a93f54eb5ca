// Image undistortion of the dataset readers, on the CPU like in the reference player:
//   EuRoC : cv::undistort(img, K, D) with the radtan model, K and D handed over as float32
//           (xrslam-pc/player/src/IO/euroc_dataset_reader.cpp:62-72)
//   TUM-VI: xrslam::extra::ImageUndistorter(..., "equidistant") + cv::remap(INTER_LINEAR, BORDER_CONSTANT)
//           (xrslam-extra/include/xrslam/extra/image_undistorter.h:13-97, tum_dataset_reader.cpp:62-76)
// The arithmetic (OpenCV's fixed-point bilinear remap) lives in xrslam_b200/csrc/undistort_math.h and is shared
// with the library's CUDA kernel (xrvio_op_undistort / xrvio_params.distortion_model), so the player needs no OpenCV
// and the CPU test of this file also covers the device code.
#pragma once
#include "../xrslam_b200/csrc/undistort_math.h"
#include "png_reader.hpp"
#include <string>

namespace xplayer {

struct Undistorter {
    int width = 0, height = 0;
    std::vector<int> qx, qy;            // source coordinates in 1/32 px

    // K = fx fy cx cy; D = k1 k2 p1 p2 (radtan) or k1 k2 k3 k4 (equidistant)
    void build(int w, int h, const double K[4], const double D[4], const std::string &model) {
        width = w; height = h;
        qx.assign((size_t)w * h, 0); qy.assign((size_t)w * h, 0);
        const xund::Model m = xund::make_model(model == "radtan" ? xund::MODEL_RADTAN : xund::MODEL_EQUIDISTANT, K, D);
        for (int v = 0; v < h; ++v)
            for (int u = 0; u < w; ++u) xund::source_q5(m, u, v, &qx[(size_t)v * w + u], &qy[(size_t)v * w + u]);
    }

    void apply(const GrayImage &src, GrayImage &dst) const {
        dst.width = width; dst.height = height; dst.data.assign((size_t)width * height, 0);
        for (size_t i = 0; i < (size_t)width * height; ++i)
            dst.data[i] = xund::remap_q5(src.data.data(), src.width, src.width, src.height, qx[i], qy[i]);
    }
};

}  // namespace xplayer
