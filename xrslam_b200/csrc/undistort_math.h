// Arithmetic of the image undistortion shared by the CUDA kernels (frontend.cu) and the CPU reader of the
// headless player (player/undistort.hpp), so that the CPU test against OpenCV covers the device code path too.
//   model 1 "radtan"      : cv::undistort(img, K, D) as called by the reference's EuRoC reader
//                           (xrslam-pc/player/src/IO/euroc_dataset_reader.cpp:62-72; K and D arrive as float32)
//   model 2 "equidistant" : xrslam::extra::ImageUndistorter (xrslam-extra/include/xrslam/extra/image_undistorter.h:13-97;
//                           the map is stored as float32 before cv::convertMaps)
// Both end in OpenCV's fixed-point bilinear remap: source coordinates rounded to 1/32 px (ties to even), weights
// (32 - fx)(32 - fy) * 32 ..., result (sum + 2^14) >> 15, zero outside the image (BORDER_CONSTANT).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define XU_HD __host__ __device__ __forceinline__
#else
#define XU_HD inline
#endif

namespace xund {

enum { MODEL_NONE = 0, MODEL_RADTAN = 1, MODEL_EQUIDISTANT = 2 };

struct Model {
    int model;
    double fx, fy, cx, cy;
    double d[4];            // radtan: k1 k2 p1 p2; equidistant: k1 k2 k3 k4
};

// the reference hands cv::undistort float32 intrinsics and coefficients; ImageUndistorter keeps doubles
XU_HD Model make_model(int model, const double K[4], const double D[4]) {
    Model m;
    m.model = model;
    const bool fl = (model == MODEL_RADTAN);
    m.fx = fl ? (double)(float)K[0] : K[0]; m.fy = fl ? (double)(float)K[1] : K[1];
    m.cx = fl ? (double)(float)K[2] : K[2]; m.cy = fl ? (double)(float)K[3] : K[3];
    for (int i = 0; i < 4; ++i) m.d[i] = fl ? (double)(float)D[i] : D[i];
    return m;
}

// source position of output pixel (u, v) in 1/32 px units
XU_HD void source_q5(const Model &m, int u, int v, int *qx, int *qy) {
    const double x = (u - m.cx) / m.fx, y = (v - m.cy) / m.fy;
    double sx, sy;
    if (m.model == MODEL_RADTAN) {
        const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
        const double kr = 1 + (m.d[1] * r2 + m.d[0]) * r2;
        const double xd = x * kr + m.d[2] * _2xy + m.d[3] * (r2 + 2 * x2);
        const double yd = y * kr + m.d[2] * (r2 + 2 * y2) + m.d[3] * _2xy;
        sx = m.fx * xd + m.cx; sy = m.fy * yd + m.cy;
    } else {
        const double r = sqrt(x * x + y * y);
        if (r < 1e-10) { sx = u; sy = v; }
        else {
            const double th = atan(r), t2 = th * th, t4 = t2 * t2, t6 = t2 * t4, t8 = t4 * t4;
            const double thd = th * (1 + m.d[0] * t2 + m.d[1] * t4 + m.d[2] * t6 + m.d[3] * t8);
            const double sc = (r > 1e-8) ? thd / r : 1.0;
            sx = m.fx * (x * sc) + m.cx; sy = m.fy * (y * sc) + m.cy;
        }
        sx = (double)(float)sx; sy = (double)(float)sy;
    }
    *qx = (int)rint(sx * 32.0);
    *qy = (int)rint(sy * 32.0);
}

XU_HD int clamp16(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

// bilinear tap at the fixed-point position (qx, qy); src is w x h with the given row stride
XU_HD unsigned char remap_q5(const unsigned char *src, int stride, int w, int h, int qx, int qy) {
    const int x = clamp16(qx >> 5), y = clamp16(qy >> 5), fx = qx & 31, fy = qy & 31;
    const bool x0 = x >= 0 && x < w, x1 = x + 1 >= 0 && x + 1 < w, y0 = y >= 0 && y < h, y1 = y + 1 >= 0 && y + 1 < h;
    const int p00 = (x0 && y0) ? src[(long)y * stride + x] : 0;
    const int p01 = (x1 && y0) ? src[(long)y * stride + x + 1] : 0;
    const int p10 = (x0 && y1) ? src[(long)(y + 1) * stride + x] : 0;
    const int p11 = (x1 && y1) ? src[(long)(y + 1) * stride + x + 1] : 0;
    const int s = p00 * (32 - fx) * (32 - fy) * 32 + p01 * fx * (32 - fy) * 32 + p10 * (32 - fx) * fy * 32 + p11 * fx * fy * 32;
    return (unsigned char)((s + (1 << 14)) >> 15);
}

}  // namespace xund
