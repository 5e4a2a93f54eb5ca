// Front-end kernels (K1-K5): CLAHE, u8 Gaussian pyramid, Harris/GFTT corner detection,
// pyramidal Lucas-Kanade tracking (forward + backward + gates), Poisson-disk spacing filter.
// Arithmetic restates what the reference obtains from OpenCV through
// xrslam-extra/src/xrslam/extra/opencv_image.cpp:38-161.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "undistort_math.h"

namespace xvfe {

constexpr int kPyrLevels = 4;        // levels 0..3 (opencv_image.h:20 level_num() == 3)
constexpr int kMaxCorners = 256;     // >= feature_tracker.max_keypoint_detection (200)
constexpr int kMaxCand = 65536;      // Harris NMS candidates kept per image
constexpr int kMaxPts = 512;         // keypoints per frame handled by KLT / Poisson

struct PyrLayout {
    int w[kPyrLevels], h[kPyrLevels];
    size_t off[kPyrLevels];
    size_t total;
};
inline PyrLayout make_pyr_layout(int w, int h) {
    PyrLayout L;
    size_t o = 0;
    for (int l = 0; l < kPyrLevels; ++l) {
        L.w[l] = w; L.h[l] = h; L.off[l] = o;
        o += (size_t)w * h;
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
    L.total = o;
    return L;
}

// ---- K1: CLAHE.  src/dst: n images of h*w bytes.  lut: n * tiles * 256 bytes scratch.
// Batch item i reads image src + (src_idx ? src_idx[i] : i) * src_stride and writes dst + (dst_idx ? ...) * dst_stride.
void clahe(const uint8_t *src, const int *src_idx, size_t src_stride, uint8_t *dst, const int *dst_idx, size_t dst_stride,
           uint8_t *lut, int n, int w, int h, double clip_limit, int tiles_x, int tiles_y, cudaStream_t st);
// ---- K2: pyramid levels 1..3 from level 0 (already in pyr + off[0]).  pyr: n * L.total bytes.
void pyramid(uint8_t *pyr, const int *idx /* pyramid index per batch item, may be null */, int n, const PyrLayout &L, cudaStream_t st);
// Scharr derivative image of one level (only for parity tests; KLT recomputes derivatives on chip).
void scharr(const uint8_t *img, int16_t *deriv, int n, int w, int h, cudaStream_t st);
// ---- K3: GFTT with Harris response.  img: n images (stride `img_stride` bytes between images).
// corners: n * kMaxCorners * 3 floats (x, y, response); count: n ints.
// scratch: cand (n * kMaxCand * 8 bytes), cand_count (n ints), maxval (n uint32)
void gftt_harris(const uint8_t *img, const int *idx, size_t img_stride, int n, int w, int h, int max_corners, float quality,
                 float min_distance, float *corners, int *count, unsigned long long *cand, int *cand_count,
                 unsigned int *maxval, cudaStream_t st);
// ---- K4: pyramidal LK forward + gates + backward + 0.5 px check (opencv_image.cpp:75-154).
// prev_pyr / next_pyr: n pyramids.  pts: n * kMaxPts * 2 floats.  npts: n ints.
// next_pts in: initial guess (gyro prediction); out: tracked position.  status out: n * kMaxPts bytes.
// maps_prev / maps_next: TMA tensor maps of the two pyramid buffers (klt_make_maps; built once per buffer), null or
// levels == 0: every tile is staged with ordinary loads.  Inner tiles are fetched by cp.async.bulk.tensor (frontend.cu).
struct alignas(64) KltMapSet { unsigned char bytes[8 * 128 + 64]; unsigned levels; };
KltMapSet klt_make_maps(const uint8_t *pyr, int n_pyramids, const PyrLayout &L);
void klt_track(const uint8_t *prev_pyr, const uint8_t *next_pyr, const int *prev_idx, const int *next_idx, int n,
               const PyrLayout &L, const float *prev_pts,
               float *next_pts, const int *npts, uint8_t *status, cudaStream_t st, const KltMapSet *maps_prev = nullptr,
               const KltMapSet *maps_next = nullptr);
// ---- K5: Poisson-disk filter.  For each of n lists: the first n_preset points are pre-seeded, the
// rest are tested in order; keep[i] = 1 if accepted.  mode 0: insert_points semantics (detect);
// mode 1: permit+preset with an extra per-point veto flag (track).  pts: n * kMaxPts * 2 doubles.
void poisson_filter(const double *pts, const int *n_preset, const int *n_total, const uint8_t *veto, uint8_t *keep,
                    int n, double radius, cudaStream_t st);

// ---- undistortion (what the reference's dataset readers do with cv::undistort / cv::remap on the CPU).
// map: w*h source positions in 1/32 px, built once per camera; apply: batch item i reads src + (src_idx ? src_idx[i] : i) *
// src_stride and writes dst + (dst_idx ? dst_idx[i] : i) * dst_stride.  1 B read-gather + 1 B write + 8 B of map (L2-resident) per pixel.
void undistort_build_map(const xund::Model &m, int w, int h, int2 *map, cudaStream_t st);
void undistort(const uint8_t *src, const int *src_idx, size_t src_stride, uint8_t *dst, const int *dst_idx, size_t dst_stride,
               int n, int w, int h, const int2 *map, cudaStream_t st);

}  // namespace xvfe
