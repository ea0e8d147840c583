// fm3d_internal.cuh -- shared declarations of libfm3d's translation units (not installed).
#ifndef FM3D_INTERNAL_CUH_
#define FM3D_INTERNAL_CUH_

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/fm3d.h"

#define FM3D_MAX_LEVELS 8
#define FM3D_SCRATCH_SLOTS 16

// ---------------------------------------------------------------- device-side camera model
// K, dist (OpenCV order k1,k2,p1,p2,k3) and g12 = [R|t] (X2 = R X1 + t), all fp64 as the
// reference keeps them (Triangulator/singlecameratriangulator.cpp:73-105,123-143).
struct fm3d_cam {
    double fx, fy, cx, cy;
    double ifx, ify;                    // 1/fx, 1/fy
    double k1, k2, p1, p2, k3;
    double R[9], t[3];
    double zmin, zmax;
};

struct fm3d_level {
    int w, h, pitch;    // pitch: bytes per row (multiple of 16 so every level is TMA-addressable)
    size_t off;         // byte offset of the level inside the pyramid allocation
};

struct fm3d_pyramid_desc {
    const uint8_t* base[2];            // image 1 / image 2
    fm3d_level lv[FM3D_MAX_LEVELS];
    int levels;                        // number of down-samplings ("pyramids" in settings.yml)
};

struct fm3d_ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    std::string err;
    // options
    int opt_geometry_f32 = 0;
    int matcher_expect_integer = 1;  // verdict of the last 128-d float matching call: 1 = launch the integer contraction without waiting for the operand check
    int opt_matcher_tensor = 1;
    int opt_matcher_persistent = 1;  // integer tensor-core matcher: one persistent CTA per SM over equal ranges of the (query tile, train tile) sequence; 0 = one CTA per (query tile, train split)
    int opt_matcher_min_tiles = 1;   // persistent matcher: fewest train tiles worth a CTA of its own (small problems use fewer CTAs)
    int opt_matcher_splits = 0;      // tensor-core matchers, one CTA per item: train splits per query tile; 0 = round 1's rule (split only below 2 CTAs per SM)
    int opt_matcher_sp_tile = 128;   // real-valued tensor filter at dim > 80: 128 = two stages of 128 train rows (default, measured faster), 256 = one stage of 256
    int opt_lm_patience = 100;
    int opt_normals_threads = 512;
    int opt_normals_tma = 1;
    int opt_normals_level_sync = 0; // two-slot kernel: the warps of a slot wait at the end of a level set-up until its LM warp has started the level (0: they go on)
    int opt_normals_pingpong = 1;  // fast kernel, four-window layout, mode 0: two features per eight warps taking turns (normals_pp_kernel)
    int opt_normals_groups = 0;    // fast kernel: feature pipelines per CTA; 0 = automatic (2 when there are more features than SMs)
    int opt_normals_memo = 3;      // fast kernel: 1 = trials whose fp32 coefficients equal the iterate's are not re-evaluated, 2 = nor are Jacobian requests (SSD), 3 = nor requests at the coefficients of one of the level's last four Jacobian passes
    int opt_normals_fuse = 3;      // fast kernel: Jacobian evaluated together with a trial: 1 = the first trial of an iteration, 2 = that, if the last one was accepted, 3 = every trial
    int opt_normals_sweep_batch = 4;  // fast kernel, dense sweep: candidates per pass (1: one pass per candidate)
    int opt_normals_fast = 1;      // 1: fm3d_normals_fast.cu (default), 0: the faithful fp64 kernel
    int opt_pyramid_fused = 1;     // K4: one fused launch per three pyramid levels (0: one pyrdown_kernel launch per level + copies)
    int opt_normals_cost = 0;      // FM3D_COST_SSD: the reference's residual I1 - I2; FM3D_COST_NCC: zero-mean normalised (fast kernel only)
    // camera
    fm3d_cam cam{};
    bool has_cam = false, has_g12 = false;
    // images
    uint8_t* pyr_mem = nullptr;
    size_t pyr_bytes = 0;
    fm3d_pyramid_desc pyr{};
    bool has_images = false;
    // grow-only scratch (device) and pinned host staging
    void* scratch[FM3D_SCRATCH_SLOTS] = {nullptr};
    size_t scratch_bytes[FM3D_SCRATCH_SLOTS] = {0};
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // counters
    int64_t n_launch = 0, n_copy = 0;
    int n_matcher_exact_fallback = 0;   // queries the last filtered float match handed to the exact path
    // an asynchronous (_dev) normal search was launched since the last fm3d_sync: its TMA-timeout flag (scratch[1][1])
    // has not been looked at yet
    bool normals_flag_pending = false;
    // several GPUs (fm3d_comm.cu): ncclComm_t of this context, its rank, collectives enqueued
    void* comm = nullptr;
    int comm_nranks = 1, comm_rank = 0;
    int64_t n_coll = 0;
};

// ---------------------------------------------------------------- error plumbing
int fm3d_fail(fm3d_ctx* ctx, int code, const char* fmt, ...);

#define FM3D_CUDA(ctx, expr)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fm3d_fail((ctx), FM3D_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                             cudaGetErrorString(e__), __FILE__, __LINE__);                     \
    } while (0)

#define FM3D_CHECK_ARG(ctx, cond)                                                              \
    do {                                                                                       \
        if (!(cond)) return fm3d_fail((ctx), FM3D_ERR_INVALID_ARG, "invalid argument: %s", #cond); \
    } while (0)

#define FM3D_LAUNCH_CHECK(ctx)                                                                 \
    do {                                                                                       \
        (ctx)->n_launch++;                                                                     \
        FM3D_CUDA((ctx), cudaGetLastError());                                                  \
    } while (0)

// Scratch slot `slot` with at least `bytes` bytes (device memory, reused across calls).
int fm3d_scratch(fm3d_ctx* ctx, int slot, size_t bytes, void** out);
int fm3d_pinned(fm3d_ctx* ctx, size_t bytes, void** out);
int fm3d_bind(fm3d_ctx* ctx);  // cudaSetDevice(ctx->device)

// Async copies on the context's stream (counted).
int fm3d_h2d(fm3d_ctx* ctx, void* dst, const void* src, size_t bytes);
int fm3d_d2h(fm3d_ctx* ctx, void* dst, const void* src, size_t bytes);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed).
int fm3d_encode_tmap_2d_u8(fm3d_ctx* ctx, CUtensorMap* map, const void* base, int w, int h,
                           int pitch, int box_w, int box_h);

// cv::SIFT's Gaussian (and DoG) pyramid in scratch slots 12 / 13 (fm3d_detect_sift.cu), shared by the detector (K14) and the
// descriptor stage for keypoints that carry an octave (K11).  Image i of octave o: gauss + gauss_off[o] + i * w[o] * h[o].
struct fm3d_sift_pyramid {
    int first_octave;                   // -1: the base image is the doubled frame
    int n_octaves, n_octave_layers;
    int w[16], h[16];
    size_t gauss_off[16], dog_off[16];  // float offsets
    size_t gauss_floats, dog_floats;
    float* gauss;
    float* dog;
};
int fm3d_sift_build_pyramid(fm3d_ctx* ctx, const uint8_t* d_img, int w, int h, int stride, int first_octave, int n_octaves_wanted,
                            int n_octave_layers, double sigma, bool with_dog, fm3d_sift_pyramid* P);
int fm3d_detect_sift_impl(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, int n_octave_layers,
                          double contrast_threshold, double edge_threshold, double sigma, int max_keypoints, float* xy, float* size,
                          float* angle, float* response, int32_t* octave, int* n, fm3d_sift_pyramid* P_keep);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

// cv::undistortPoints without R/P: 5 fixed-point iterations in fp64
// (call sites Triangulator/singlecameratriangulator.cpp:169-170,542).
__device__ __forceinline__ void fm3d_undistort(const fm3d_cam& c, double u, double v, double& xo,
                                               double& yo) {
    const double x0 = (u - c.cx) / c.fx, y0 = (v - c.cy) / c.fy;
    double x = x0, y = y0;
#pragma unroll 1
    for (int it = 0; it < 5; it++) {
        const double r2 = x * x + y * y;
        const double icd = 1.0 / (1.0 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
        if (icd < 0) { x = x0; y = y0; break; }
        const double dx = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x);
        const double dy = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
        x = (x0 - dx) * icd;
        y = (y0 - dy) * icd;
    }
    xo = x; yo = y;
}

// cv::projectPoints on a point already in the camera frame: perspective division, lens
// distortion, K (call sites singlecameratriangulator.cpp:388,602,735,817).
template <typename T>
__device__ __forceinline__ void fm3d_distort_K(T x, T y, T k1, T k2, T p1, T p2, T k3, T fx, T fy,
                                               T cx, T cy, T& u, T& v) {
    const T r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const T a1 = T(2) * x * y, a2 = r2 + T(2) * x * x, a3 = r2 + T(2) * y * y;
    const T cdist = T(1) + k1 * r2 + k2 * r4 + k3 * r6;
    const T xd = x * cdist + p1 * a1 + p2 * a2;
    const T yd = y * cdist + p1 * a3 + p2 * a1;
    u = xd * fx + cx;
    v = yd * fy + cy;
}

__device__ __forceinline__ void fm3d_project(const fm3d_cam& c, double X, double Y, double Z,
                                             double& u, double& v) {
    const double z = Z != 0.0 ? 1.0 / Z : 1.0;
    fm3d_distort_K<double>(X * z, Y * z, c.k1, c.k2, c.p1, c.p2, c.k3, c.fx, c.fy, c.cx, c.cy, u, v);
}

// isPixelGood (singlecameratriangulator.cpp:657-665): inv_scale = 1/scale, cols/rows of the
// current pyramid level.
__device__ __forceinline__ bool fm3d_pixel_good(double x, double y, double inv_scale, int cols,
                                                int rows) {
    return !((x < 0) || (x > inv_scale * cols) || (y < 0) || (y > inv_scale * rows));
}

// The float arithmetic of getBilinearInterpPix32f (tools.cpp:129-142): every multiply and add
// rounded separately, in the reference's expression order.
__device__ __forceinline__ float fm3d_lerp4(float b00, float b01, float b10, float b11, float ax,
                                            float ay) {
    const float xm0 = __fsub_rn(1.0f, ax), ym0 = __fsub_rn(1.0f, ay);
    const float c0 = __fadd_rn(__fmul_rn(b00, ym0), __fmul_rn(b01, ay));
    const float c1 = __fadd_rn(__fmul_rn(b10, ym0), __fmul_rn(b11, ay));
    return __fadd_rn(__fmul_rn(xm0, c0), __fmul_rn(ax, c1));
}

// u8 -> float without the quarter-rate I2F: 0x4B000000 | b is 8388608 + b exactly.
__device__ __forceinline__ float fm3d_u8f(unsigned b) {
    return __uint_as_float(0x4B000000u | b) - 8388608.0f;
}

// at<uchar>(y, x) of a continuous w x h image stored with row pitch `pitch`: flat addressing
// y*w + x like cv::Mat::at without bounds checks; bytes outside the buffer read as 0 (D2).
__device__ __forceinline__ unsigned fm3d_at_flat(const uint8_t* __restrict__ img, int w, int h,
                                                 int pitch, int x, int y) {
    if (x >= 0 && x < w && y >= 0 && y < h) return img[(size_t)y * pitch + x];
    const long long idx = (long long)y * w + x;
    if (idx < 0 || idx >= (long long)w * h) return 0u;
    const int yy = (int)(idx / w), xx = (int)(idx - (long long)yy * w);
    return img[(size_t)yy * pitch + xx];
}

// getBilinearInterpPix32f on a pitched global-memory image.
__device__ __forceinline__ float fm3d_bilinear_global(const uint8_t* __restrict__ img, int w, int h,
                                                      int pitch, float x, float y) {
    const float fx0 = floorf(x), fy0 = floorf(y);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float b00 = fm3d_u8f(fm3d_at_flat(img, w, h, pitch, x0, y0));
    const float b01 = fm3d_u8f(fm3d_at_flat(img, w, h, pitch, x0, y0 + 1));
    const float b10 = fm3d_u8f(fm3d_at_flat(img, w, h, pitch, x0 + 1, y0));
    const float b11 = fm3d_u8f(fm3d_at_flat(img, w, h, pitch, x0 + 1, y0 + 1));
    return fm3d_lerp4(b00, b01, b10, b11, __fsub_rn(x, fx0), __fsub_rn(y, fy0));
}

#endif  // __CUDACC__
#endif  // FM3D_INTERNAL_CUH_
