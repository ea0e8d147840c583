// fm3d_detect_sift.cu -- K14: cv::SIFT's scale-space detector on the GPU (fm3d_detect_sift), and the Gaussian pyramid it
// shares with the descriptor stage (fm3d_describe_keypoints_sift for keypoints that carry an octave).
//
// Replaces feature_detector_->detect(frame, keypoints) of DescriptorsMatcher::compareWithNNDR / compare / crosscompare
// (DescriptorsMatcher/descriptorsmatcher.cpp:110-111, :91-92, :76-77) for DetectorType SIFT (:243-256: cv::SIFT(NumFeatures,
// NumOctaveLayers, ContrastThreshold, EdgeThreshold, Sigma)).  The algorithm is OpenCV's (third party, unpinned in the
// reference; restated in oracle/sift_detect_np.py, which is pinned to cv2.SIFT_create().detect):
//   createInitialImage     float(gray) doubled (INTER_LINEAR: exact for 8-bit values), blurred with sqrt(sigma^2 - 1)
//   buildGaussianPyramid   cvRound(log2(min side of the doubled image) - 2) + 1 octaves of nOctaveLayers + 3 images
//   buildDoGPyramid        differences of neighbouring images
//   findScaleSpaceExtrema  26-neighbour extrema above the contrast floor, adjustLocalExtrema (<= 5 Newton steps of the 3-D
//                          quadratic fit, closed-form 3 x 3 solve, contrast and edge tests), calcOrientationHist (36 bins,
//                          (1 4 6 4 1)/16 smoothing), one keypoint per peak >= 0.8 max
//   removeDuplicatedSorted / retainBest / back to the original image scale
// Kernels: sift_up2_kernel, sift_blur_rows_kernel / sift_blur_cols_kernel (separable, <= 27 taps, BORDER_REFLECT_101),
// sift_half_kernel, sift_dog_kernel, sift_extrema_kernel (one thread per pixel, candidates appended to a list),
// sift_refine_kernel (one warp per candidate: lane 0 runs the Newton steps, the warp builds the orientation histogram in
// per-lane private bins -- fixed summation order, no atomics).  The keypoints are sorted / de-duplicated on the host in
// OpenCV's order, so the result does not depend on the order the warps finish in.
// Float sums are taken in this file's order, not OpenCV's SIMD order: values agree to a few ulp, a candidate within rounding
// of a threshold (contrast, edge, 0.8 peak, |offset| = 0.5) can fall on the other side (tests/test_gpu_sift_detect.py gates
// the matched fraction).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <functional>
#include <vector>

#include "fm3d_internal.cuh"

namespace {

constexpr int SIFT_IMG_BORDER = 5;
constexpr int SIFT_MAX_INTERP_STEPS = 5;
constexpr int SIFT_ORI_HIST_BINS = 36;
constexpr float SIFT_ORI_SIG_FCTR = 1.5f;
constexpr float SIFT_ORI_RADIUS = 4.5f;
constexpr float SIFT_ORI_PEAK_RATIO = 0.8f;
constexpr int MAX_TAPS = 16;            // half width + 1 of the widest kernel (sigma 3.09 * 8 + 1 = 27 taps -> 14)
constexpr int MAX_LAYERS = 8;           // nOctaveLayers + 3 <= 8

struct BlurKernel {
    int radius;
    float k[MAX_TAPS];                  // k[0] centre, k[j] = weight of +-j
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// createInitialImage: dst(2W x 2H) = resize(float(src), INTER_LINEAR): the source coordinate of dx is (dx + 0.5) / 2 - 0.5,
// i.e. taps sx = floor, weights 0.75 / 0.25, clamped at the border (cv::resize replicates).  Exact for 8-bit values.
__global__ void sift_up2_kernel(const uint8_t* __restrict__ src, int w, int h, int stride, float* __restrict__ dst) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y * blockDim.y + threadIdx.y;
    if (dx >= 2 * w || dy >= 2 * h) return;
    const float fx = (dx + 0.5f) * 0.5f - 0.5f, fy = (dy + 0.5f) * 0.5f - 0.5f;
    int sx = (int)floorf(fx), sy = (int)floorf(fy);
    float ax = fx - (float)sx, ay = fy - (float)sy;
    if (sx < 0) { sx = 0; ax = 0.f; }
    if (sx >= w - 1) { sx = w - 1; ax = 0.f; }
    if (sy < 0) { sy = 0; ay = 0.f; }
    if (sy >= h - 1) { sy = h - 1; ay = 0.f; }
    const int sx1 = min(sx + 1, w - 1), sy1 = min(sy + 1, h - 1);
    const float a = src[(size_t)sy * stride + sx], b = src[(size_t)sy * stride + sx1];
    const float c = src[(size_t)sy1 * stride + sx], d = src[(size_t)sy1 * stride + sx1];
    const float r0 = a * (1.f - ax) + b * ax, r1 = c * (1.f - ax) + d * ax;
    dst[(size_t)dy * (2 * w) + dx] = r0 * (1.f - ay) + r1 * ay;
}
// firstOctave = 0 (descriptor stage for keypoints of octave >= 0 only): float(gray)
__global__ void sift_u8f_kernel(const uint8_t* __restrict__ src, int w, int h, int stride, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < w && y < h) dst[(size_t)y * w + x] = (float)src[(size_t)y * stride + x];
}

__global__ void sift_blur_rows_kernel(const float* __restrict__ src, int w, int h, BlurKernel K, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const float* row = src + (size_t)y * w;
    float s = K.k[0] * row[x];
    if (x >= K.radius && x + K.radius < w) {
        for (int j = 1; j <= K.radius; j++) s = fmaf(K.k[j], row[x - j] + row[x + j], s);
    } else {
        for (int j = 1; j <= K.radius; j++) s = fmaf(K.k[j], row[reflect101(x - j, w)] + row[reflect101(x + j, w)], s);
    }
    dst[(size_t)y * w + x] = s;
}
__global__ void sift_blur_cols_kernel(const float* __restrict__ src, int w, int h, BlurKernel K, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    float s = K.k[0] * src[(size_t)y * w + x];
    if (y >= K.radius && y + K.radius < h) {
        for (int j = 1; j <= K.radius; j++) s = fmaf(K.k[j], src[(size_t)(y - j) * w + x] + src[(size_t)(y + j) * w + x], s);
    } else {
        for (int j = 1; j <= K.radius; j++)
            s = fmaf(K.k[j], src[(size_t)reflect101(y - j, h) * w + x] + src[(size_t)reflect101(y + j, h) * w + x], s);
    }
    dst[(size_t)y * w + x] = s;
}
// Both passes of a Gaussian blur for a 64 x 32 tile in shared memory (what sift_blur_rows_kernel + sift_blur_cols_kernel do through
// a scratch image: the same sums in the same order, so the same bits): the tile and its halo of `radius` pixels are staged with
// BORDER_REFLECT_101, the row pass runs over the tile's rows and halo rows, the column pass over the tile.  With `prev` / `dog`
// the difference to the previous image of the octave (buildDoGPyramid) is written on the way out.
constexpr int BT_W = 64, BT_H = 32, BT_R = MAX_TAPS - 1;
__global__ void __launch_bounds__(256)
sift_blur_tile_kernel(const float* __restrict__ src, int w, int h, BlurKernel K, float* __restrict__ dst,
                      const float* __restrict__ prev, float* __restrict__ dog) {
    __shared__ float in[BT_H + 2 * BT_R][BT_W + 2 * BT_R + 1];
    __shared__ float mid[BT_H + 2 * BT_R][BT_W + 1];
    const int R = K.radius, tid = threadIdx.x;
    const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H;
    const int iw = BT_W + 2 * R, ih = BT_H + 2 * R;
    for (int i = tid; i < iw * ih; i += 256) {
        const int ry = i / iw, rx = i - ry * iw;
        in[ry][rx] = src[(size_t)reflect101(y0 - R + ry, h) * w + reflect101(x0 - R + rx, w)];
    }
    __syncthreads();
    for (int i = tid; i < ih * BT_W; i += 256) {
        const int ry = i / BT_W, x = i - ry * BT_W;
        const float* row = &in[ry][x + R];
        float s = K.k[0] * row[0];
        for (int j = 1; j <= R; j++) s = fmaf(K.k[j], row[-j] + row[j], s);
        mid[ry][x] = s;
    }
    __syncthreads();
    for (int i = tid; i < BT_H * BT_W; i += 256) {
        const int y = i / BT_W, x = i - y * BT_W;
        const int gx = x0 + x, gy = y0 + y;
        if (gx >= w || gy >= h) continue;
        float s = K.k[0] * mid[y + R][x];
        for (int j = 1; j <= R; j++) s = fmaf(K.k[j], mid[y + R - j][x] + mid[y + R + j][x], s);
        const size_t o = (size_t)gy * w + gx;
        dst[o] = s;
        if (dog) dog[o] = s - prev[o];
    }
}
// first image of the next octave: every second pixel (cv::resize INTER_NEAREST to half size)
__global__ void sift_half_kernel(const float* __restrict__ src, int w, int h, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    const int w2 = w / 2;
    if (x < w2) dst[(size_t)y * w2 + x] = src[(size_t)(2 * y) * w + 2 * x];
}
__global__ void sift_dog_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, float* __restrict__ d) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = b[i] - a[i];
}

struct Cand { int octave, layer, r, c; };

// findScaleSpaceExtrema, the per-pixel test (sift.simd.hpp: |val| > threshold and val >= / <= its 26 neighbours)
__global__ void sift_extrema_kernel(const float* __restrict__ dog, int w, int h, int octave, int n_layers, float threshold,
                                    Cand* __restrict__ cand, int* __restrict__ n_cand, int cap) {
    const int c = SIFT_IMG_BORDER + blockIdx.x * blockDim.x + threadIdx.x, r = SIFT_IMG_BORDER + blockIdx.y;
    const int layer = 1 + blockIdx.z;
    if (c >= w - SIFT_IMG_BORDER || r >= h - SIFT_IMG_BORDER) return;
    const size_t plane = (size_t)w * h;
    const float* cur = dog + plane * layer;
    const float val = cur[(size_t)r * w + c];
    if (!(fabsf(val) > threshold)) return;
    bool is_max = val > 0.f, is_min = val < 0.f;
    for (int s = -1; s <= 1 && (is_max || is_min); s++) {
        const float* p = cur + (ptrdiff_t)plane * s + (size_t)r * w + c;
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                const float v = p[dy * w + dx];
                is_max = is_max && val >= v;
                is_min = is_min && val <= v;
            }
    }
    if (is_max || is_min) {
        const int k = atomicAdd(n_cand, 1);
        if (k < cap) cand[k] = Cand{octave, layer, r, c};
    }
    (void)n_layers;
}

struct Octave {
    size_t gauss_off, dog_off;          // float offsets of the octave's first image in the two pyramids
    int w, h;
};
struct RefineArgs {
    const float* gauss;
    const float* dog;
    Octave oct[16];
    int n_octave_layers;
    float contrast_threshold, edge_threshold, sigma;
    const Cand* cand;
    int n_cand;
    float* out;                          // 6 floats per keypoint: x, y, size, angle, response, octave (int bits)
    int* n_out;
    int cap;
};

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {    // cv::fastAtan2
    const float s = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + 2.220446049250313e-16f);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + 2.220446049250313e-16f);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// Matx33f::solve(b, DECOMP_LU) for 3 x 3: the closed form of Matx_FastSolveOp<float, 3, 3, 1>
__device__ __forceinline__ bool solve3(const float a[3][3], const float b[3], float x[3]) {
    float d = a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) - a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
              a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]);
    if (d == 0.f) return false;
    d = 1.f / d;
    x[0] = d * (b[0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (b[1] * a[2][2] - a[1][2] * b[2]) +
                a[0][2] * (b[1] * a[2][1] - a[1][1] * b[2]));
    x[1] = d * (a[0][0] * (b[1] * a[2][2] - a[1][2] * b[2]) - b[0] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                a[0][2] * (a[1][0] * b[2] - b[1] * a[2][0]));
    x[2] = d * (a[0][0] * (a[1][1] * b[2] - b[1] * a[2][1]) - a[0][1] * (a[1][0] * b[2] - b[1] * a[2][0]) +
                b[0] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]));
    return true;
}

constexpr int RF_WARPS = 4;

__global__ void __launch_bounds__(RF_WARPS * 32)
sift_refine_kernel(const RefineArgs A) {
    __shared__ float lanehist[RF_WARPS][SIFT_ORI_HIST_BINS][32];     // per-lane private bins: bank = lane
    __shared__ float hist_s[RF_WARPS][SIFT_ORI_HIST_BINS + 4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * RF_WARPS + warp;
    if (ci >= A.n_cand) return;
    const Cand cd = A.cand[ci];
    const Octave O = A.oct[cd.octave];
    const int w = O.w, h = O.h, nol = A.n_octave_layers;
    const size_t plane = (size_t)w * h;
    const float* dog = A.dog + O.dog_off;

    // ---- adjustLocalExtrema on lane 0
    int ok = 0, r = cd.r, c = cd.c, layer = cd.layer;
    float xi = 0.f, xr = 0.f, xc = 0.f, contr = 0.f;
    if (lane == 0) {
        const float img_scale = 1.f / 255.f, deriv_scale = img_scale * 0.5f, second_deriv_scale = img_scale, cross_deriv_scale = img_scale * 0.25f;
        int i = 0;
        bool alive = true;
        for (; i < SIFT_MAX_INTERP_STEPS; i++) {
            const float* img = dog + plane * layer + (size_t)r * w + c;
            const float* prev = img - plane;
            const float* next = img + plane;
            const float dD[3] = {(img[1] - img[-1]) * deriv_scale, (img[w] - img[-w]) * deriv_scale, (next[0] - prev[0]) * deriv_scale};
            const float v2 = img[0] * 2.f;
            const float dxx = (img[1] + img[-1] - v2) * second_deriv_scale;
            const float dyy = (img[w] + img[-w] - v2) * second_deriv_scale;
            const float dss = (next[0] + prev[0] - v2) * second_deriv_scale;
            const float dxy = (img[w + 1] - img[w - 1] - img[-w + 1] + img[-w - 1]) * cross_deriv_scale;
            const float dxs = (next[1] - next[-1] - prev[1] + prev[-1]) * cross_deriv_scale;
            const float dys = (next[w] - next[-w] - prev[w] + prev[-w]) * cross_deriv_scale;
            const float Hm[3][3] = {{dxx, dxy, dxs}, {dxy, dyy, dys}, {dxs, dys, dss}};
            float X[3] = {0.f, 0.f, 0.f};
            solve3(Hm, dD, X);
            xi = -X[2]; xr = -X[1]; xc = -X[0];
            if (fabsf(xi) < 0.5f && fabsf(xr) < 0.5f && fabsf(xc) < 0.5f) break;
            const float big = (float)(INT_MAX / 3);
            if (!(fabsf(xi) <= big && fabsf(xr) <= big && fabsf(xc) <= big)) { alive = false; break; }
            c += __float2int_rn(xc); r += __float2int_rn(xr); layer += __float2int_rn(xi);
            if (layer < 1 || layer > nol || c < SIFT_IMG_BORDER || c >= w - SIFT_IMG_BORDER || r < SIFT_IMG_BORDER || r >= h - SIFT_IMG_BORDER) {
                alive = false;
                break;
            }
        }
        if (alive && i < SIFT_MAX_INTERP_STEPS) {
            const float* img = dog + plane * layer + (size_t)r * w + c;
            const float* prev = img - plane;
            const float* next = img + plane;
            const float dD[3] = {(img[1] - img[-1]) * deriv_scale, (img[w] - img[-w]) * deriv_scale, (next[0] - prev[0]) * deriv_scale};
            const float t = dD[0] * xc + dD[1] * xr + dD[2] * xi;
            contr = img[0] * img_scale + t * 0.5f;
            if (!(fabsf(contr) * nol < A.contrast_threshold)) {
                const float v2 = img[0] * 2.f;
                const float dxx = (img[1] + img[-1] - v2) * second_deriv_scale;
                const float dyy = (img[w] + img[-w] - v2) * second_deriv_scale;
                const float dxy = (img[w + 1] - img[w - 1] - img[-w + 1] + img[-w - 1]) * cross_deriv_scale;
                const float tr = dxx + dyy, det = dxx * dyy - dxy * dxy;
                const float e = A.edge_threshold;
                if (!(det <= 0.f || tr * tr * e >= (e + 1.f) * (e + 1.f) * det)) ok = 1;
            }
        }
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (!ok) return;
    r = __shfl_sync(0xffffffffu, r, 0); c = __shfl_sync(0xffffffffu, c, 0); layer = __shfl_sync(0xffffffffu, layer, 0);
    xi = __shfl_sync(0xffffffffu, xi, 0); xr = __shfl_sync(0xffffffffu, xr, 0); xc = __shfl_sync(0xffffffffu, xc, 0);
    contr = __shfl_sync(0xffffffffu, contr, 0);
    const int octv = cd.octave;
    const float oscale = (float)(1 << octv);
    const float kx = (c + xc) * oscale, ky = (r + xr) * oscale;
    const int koct = octv + (layer << 8) + (__float2int_rn((xi + 0.5f) * 255.f) << 16);
    const float ksize = A.sigma * powf(2.f, (layer + xi) / nol) * oscale * 2.f;
    const float scl_octv = ksize * 0.5f / oscale;

    // ---- calcOrientationHist on the Gaussian image of the keypoint's layer
    const float* gimg = A.gauss + O.gauss_off + plane * layer;
    const int radius = __float2int_rn(SIFT_ORI_RADIUS * scl_octv);
    const float sg = SIFT_ORI_SIG_FCTR * scl_octv;
    const float expf_scale = -1.f / (2.f * sg * sg);
    float* my = &lanehist[warp][0][lane];
#pragma unroll
    for (int b = 0; b < SIFT_ORI_HIST_BINS; b++) my[b * 32] = 0.f;
    const int side = 2 * radius + 1;
    for (int k = lane; k < side * side; k += 32) {
        const int i = k / side - radius, j = k - (k / side) * side - radius;
        const int y = r + i, x = c + j;
        if (y <= 0 || y >= h - 1 || x <= 0 || x >= w - 1) continue;
        const float* p = gimg + (size_t)y * w + x;
        const float dx = p[1] - p[-1], dy = p[-w] - p[w];
        const float wgt = expf((float)(i * i + j * j) * expf_scale);
        const float ori = fast_atan2_deg(dy, dx);
        const float mag = sqrtf(dx * dx + dy * dy);
        int bin = __float2int_rn((SIFT_ORI_HIST_BINS / 360.f) * ori);
        if (bin >= SIFT_ORI_HIST_BINS) bin -= SIFT_ORI_HIST_BINS;
        if (bin < 0) bin += SIFT_ORI_HIST_BINS;
        my[bin * 32] += wgt * mag;
    }
    __syncwarp();
    // fixed-order sum over the lanes, two bins per lane
    for (int b = lane; b < SIFT_ORI_HIST_BINS; b += 32) {
        float s = 0.f;
        for (int l = 0; l < 32; l++) s += lanehist[warp][b][l];
        hist_s[warp][b + 2] = s;
    }
    __syncwarp();
    if (lane == 0) {
        float* t = &hist_s[warp][0];
        const int n = SIFT_ORI_HIST_BINS;
        t[0] = t[n]; t[1] = t[n + 1]; t[n + 2] = t[2]; t[n + 3] = t[3];
        float hist[SIFT_ORI_HIST_BINS];
        float omax = 0.f;
        for (int i = 0; i < n; i++) {
            hist[i] = (t[i] + t[i + 4]) * (1.f / 16.f) + (t[i + 1] + t[i + 3]) * (4.f / 16.f) + t[i + 2] * (6.f / 16.f);
            omax = i == 0 ? hist[0] : fmaxf(omax, hist[i]);
        }
        const float mag_thr = omax * SIFT_ORI_PEAK_RATIO;
        for (int j = 0; j < n; j++) {
            const int l = j > 0 ? j - 1 : n - 1, r2 = j < n - 1 ? j + 1 : 0;
            if (hist[j] > hist[l] && hist[j] > hist[r2] && hist[j] >= mag_thr) {
                float bin = j + 0.5f * (hist[l] - hist[r2]) / (hist[l] - 2.f * hist[j] + hist[r2]);
                bin = bin < 0 ? n + bin : (bin >= n ? bin - n : bin);
                float angle = 360.f - (360.f / n) * bin;
                if (fabsf(angle - 360.f) < 1.1920929e-07f) angle = 0.f;
                const int k = atomicAdd(A.n_out, 1);
                if (k < A.cap) {
                    float* o = A.out + 6 * (size_t)k;
                    o[0] = kx; o[1] = ky; o[2] = ksize; o[3] = angle; o[4] = fabsf(contr); o[5] = __int_as_float(koct);
                }
            }
        }
    }
}

// getGaussianKernel(ksize, sigma, CV_32F) with ksize = cvRound(sigma * 8 + 1) | 1 (the float branch of GaussianBlur)
BlurKernel gaussian_kernel(double sigma) {
    BlurKernel K;
    int ksize = (int)std::nearbyint(sigma * 8 + 1) | 1;
    if (ksize > 2 * (MAX_TAPS - 1) + 1) ksize = 2 * (MAX_TAPS - 1) + 1;
    const int R = ksize / 2;
    std::vector<float> cf(ksize);
    const double scale2X = -0.5 / (sigma * sigma);
    double sum = 0;
    for (int i = 0; i < ksize; i++) {
        const double x = i - (ksize - 1) * 0.5;
        cf[i] = (float)std::exp(scale2X * x * x);
        sum += cf[i];
    }
    sum = 1. / sum;
    K.radius = R;
    for (int j = 0; j < MAX_TAPS; j++) K.k[j] = 0.f;
    for (int j = 0; j <= R; j++) K.k[j] = (float)(cf[R + j] * sum);
    return K;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// Pyramid plan + build, shared with the descriptor stage (fm3d_describe_kp.cu).
int fm3d_sift_build_pyramid(fm3d_ctx* ctx, const uint8_t* d_img, int w, int h, int stride, int first_octave, int n_octaves_wanted,
                            int n_octave_layers, double sigma, bool with_dog, fm3d_sift_pyramid* P) {
    const int L = n_octave_layers + 3;
    if (L > MAX_LAYERS) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "SIFT: nOctaveLayers > %d", MAX_LAYERS - 3);
    const int bw = first_octave < 0 ? 2 * w : w, bh = first_octave < 0 ? 2 * h : h;
    int n_oct = n_octaves_wanted > 0 ? n_octaves_wanted
                                     : (int)std::nearbyint(std::log((double)std::min(bw, bh)) / std::log(2.) - 2) - first_octave;
    if (n_oct < 1) n_oct = 1;
    if (n_oct > 16) n_oct = 16;
    P->first_octave = first_octave; P->n_octave_layers = n_octave_layers;
    size_t go = 0, dg = 0;
    int ow = bw, oh = bh, no = 0;
    for (int o = 0; o < n_oct; o++) {
        if (ow < 1 || oh < 1) break;
        P->w[o] = ow; P->h[o] = oh;
        P->gauss_off[o] = go; P->dog_off[o] = dg;
        go += (size_t)ow * oh * L;
        dg += (size_t)ow * oh * (L - 1);
        ow /= 2; oh /= 2;
        no++;
    }
    P->n_octaves = no;
    P->gauss_floats = go; P->dog_floats = dg;
    float* gauss = nullptr;
    float* dog = nullptr;
    float* tmp = nullptr;
    if (int rc = fm3d_scratch(ctx, 12, sizeof(float) * go, (void**)&gauss)) return rc;
    if (with_dog) if (int rc = fm3d_scratch(ctx, 13, sizeof(float) * dg, (void**)&dog)) return rc;
    if (int rc = fm3d_scratch(ctx, 14, sizeof(float) * 2 * (size_t)bw * bh, (void**)&tmp)) return rc;
    P->gauss = gauss; P->dog = dog;

    // createInitialImage
    const dim3 b2(32, 8);
    float* raw = tmp + (size_t)bw * bh;
    if (first_octave < 0) {
        sift_up2_kernel<<<dim3((bw + 31) / 32, (bh + 7) / 8), b2, 0, ctx->stream>>>(d_img, w, h, stride, raw);
    } else {
        sift_u8f_kernel<<<dim3((bw + 31) / 32, (bh + 7) / 8), b2, 0, ctx->stream>>>(d_img, w, h, stride, raw);
    }
    FM3D_LAUNCH_CHECK(ctx);
    const float init_sigma = 0.5f;
    const float s = (float)sigma;
    const float sig_diff = first_octave < 0 ? sqrtf(std::max(s * s - init_sigma * init_sigma * 4, 0.01f))
                                            : sqrtf(std::max(s * s - init_sigma * init_sigma, 0.01f));
    // prev / dogp: also write dogp = dst - prev (the DoG image between this layer and the previous one)
    auto blur = [&](const float* src, float* dst, int ww, int hh, double sg, const float* prev, float* dogp) -> int {
        const BlurKernel K = gaussian_kernel(sg);
        if (ww > 2 * K.radius && hh > 2 * K.radius) {
            sift_blur_tile_kernel<<<dim3((ww + BT_W - 1) / BT_W, (hh + BT_H - 1) / BT_H), 256, 0, ctx->stream>>>(src, ww, hh, K, dst, prev, dogp);
            FM3D_LAUNCH_CHECK(ctx);
            return FM3D_OK;
        }
        sift_blur_rows_kernel<<<dim3((ww + 127) / 128, hh), 128, 0, ctx->stream>>>(src, ww, hh, K, tmp);
        FM3D_LAUNCH_CHECK(ctx);
        sift_blur_cols_kernel<<<dim3((ww + 127) / 128, hh), 128, 0, ctx->stream>>>(tmp, ww, hh, K, dst);
        FM3D_LAUNCH_CHECK(ctx);
        if (dogp) {
            const size_t n = (size_t)ww * hh;
            sift_dog_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(prev, dst, n, dogp);
            FM3D_LAUNCH_CHECK(ctx);
        }
        return FM3D_OK;
    };
    if (int rc = blur(raw, gauss, bw, bh, (double)sig_diff, nullptr, nullptr)) return rc;
    // buildGaussianPyramid
    double sig[MAX_LAYERS];
    sig[0] = sigma;
    const double k = std::pow(2., 1. / n_octave_layers);
    for (int i = 1; i < L; i++) {
        const double sig_prev = std::pow(k, (double)(i - 1)) * sigma, sig_total = sig_prev * k;
        sig[i] = std::sqrt(sig_total * sig_total - sig_prev * sig_prev);
    }
    for (int o = 0; o < no; o++) {
        const int ww = P->w[o], hh = P->h[o];
        const size_t plane = (size_t)ww * hh;
        float* oct = gauss + P->gauss_off[o];
        if (o > 0) {
            const float* src = gauss + P->gauss_off[o - 1] + (size_t)P->w[o - 1] * P->h[o - 1] * n_octave_layers;
            sift_half_kernel<<<dim3((ww + 127) / 128, hh), 128, 0, ctx->stream>>>(src, P->w[o - 1], P->h[o - 1], oct);
            FM3D_LAUNCH_CHECK(ctx);
        }
        for (int i = 1; i < L; i++)
            if (int rc = blur(oct + plane * (i - 1), oct + plane * i, ww, hh, sig[i], with_dog ? oct + plane * (i - 1) : nullptr,
                              with_dog ? dog + P->dog_off[o] + plane * (i - 1) : nullptr)) return rc;
    }
    return FM3D_OK;
}

// cv::SIFT::detect.  P_keep != nullptr: the pyramid the keypoints were found on (first octave -1, Gaussian images in scratch
// 12) is described there and stays valid until the next call that builds one: fm3d_detect_and_describe_sift describes the
// keypoints on it instead of building it a second time.
int fm3d_detect_sift_impl(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, int n_octave_layers,
                          double contrast_threshold, double edge_threshold, double sigma, int max_keypoints, float* xy, float* size,
                          float* angle, float* response, int32_t* octave, int* n, fm3d_sift_pyramid* P_keep) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, img && n && w >= 2 && h >= 2 && stride >= w && n_octave_layers >= 1 && sigma > 0 && max_keypoints >= 0);
    FM3D_CHECK_ARG(ctx, max_keypoints == 0 || (xy && size && angle && response && octave));
    if (int rc = fm3d_bind(ctx)) return rc;
    uint8_t* d_img = nullptr;
    if (int rc = fm3d_scratch(ctx, 15, (size_t)w * h, (void**)&d_img)) return rc;
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(d_img, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    fm3d_sift_pyramid P;
    if (int rc = fm3d_sift_build_pyramid(ctx, d_img, w, h, w, -1, 0, n_octave_layers, sigma, true, &P)) return rc;
    if (P_keep) *P_keep = P;

    // candidate and keypoint lists (scratch 0): counters, candidates, keypoints
    const int cap_c = 1 << 20, cap_k = 1 << 20;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 256 + sizeof(Cand) * (size_t)cap_c + sizeof(float) * 6 * (size_t)cap_k, (void**)&d)) return rc;
    int* counters = reinterpret_cast<int*>(d);
    Cand* cand = reinterpret_cast<Cand*>(d + 256);
    float* out = reinterpret_cast<float*>(d + 256 + sizeof(Cand) * (size_t)cap_c);
    FM3D_CUDA(ctx, cudaMemsetAsync(counters, 0, 256, ctx->stream));
    ctx->n_copy++;
    const float threshold = (float)std::floor(0.5 * contrast_threshold / n_octave_layers * 255);
    for (int o = 0; o < P.n_octaves; o++) {
        const int ww = P.w[o], hh = P.h[o];
        if (ww <= 2 * SIFT_IMG_BORDER || hh <= 2 * SIFT_IMG_BORDER) continue;
        dim3 grid((ww - 2 * SIFT_IMG_BORDER + 127) / 128, hh - 2 * SIFT_IMG_BORDER, n_octave_layers);
        sift_extrema_kernel<<<grid, 128, 0, ctx->stream>>>(P.dog + P.dog_off[o], ww, hh, o, n_octave_layers, threshold, cand, counters, cap_c);
        FM3D_LAUNCH_CHECK(ctx);
    }
    int h_counts[2] = {0, 0};
    if (int rc = fm3d_d2h(ctx, h_counts, counters, sizeof(int))) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    int n_cand = std::min(h_counts[0], cap_c);
    std::vector<float> K;
    int nk = 0;
    if (n_cand > 0) {
        RefineArgs A;
        A.gauss = P.gauss; A.dog = P.dog;
        for (int o = 0; o < P.n_octaves; o++) A.oct[o] = Octave{P.gauss_off[o], P.dog_off[o], P.w[o], P.h[o]};
        A.n_octave_layers = n_octave_layers;
        A.contrast_threshold = (float)contrast_threshold; A.edge_threshold = (float)edge_threshold; A.sigma = (float)sigma;
        A.cand = cand; A.n_cand = n_cand; A.out = out; A.n_out = counters + 1; A.cap = cap_k;
        sift_refine_kernel<<<(n_cand + RF_WARPS - 1) / RF_WARPS, RF_WARPS * 32, 0, ctx->stream>>>(A);
        FM3D_LAUNCH_CHECK(ctx);
        if (int rc = fm3d_d2h(ctx, h_counts + 1, counters + 1, sizeof(int))) return rc;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        nk = std::min(h_counts[1], cap_k);
        K.resize(6 * (size_t)nk);
        if (nk > 0) {
            if (int rc = fm3d_d2h(ctx, K.data(), out, sizeof(float) * 6 * (size_t)nk)) return rc;
            FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    // KeyPointsFilter::removeDuplicatedSorted, retainBest, back to the original image (firstOctave = -1)
    struct Kp { float x, y, size, angle, response; int octave; };
    std::vector<Kp> kp(nk);
    for (int i = 0; i < nk; i++) {
        const float* p = &K[6 * (size_t)i];
        int oc;
        memcpy(&oc, p + 5, 4);
        kp[i] = Kp{p[0], p[1], p[2], p[3], p[4], oc};
    }
    // KeyPoint12_LessThan: x, y ascending, then size descending, angle ascending, response descending, octave descending.
    // x and y are non-negative floats (their bit patterns order like the values): an LSD radix sort on the 64-bit key (x, y),
    // then the full comparison inside the short runs of equal (x, y) (the orientations of one extremum) -- 129 k keypoints of
    // a 4K frame in ~1 ms where std::sort with the six-way comparator took ~10 ms.
    auto less = [](const Kp& a, const Kp& b) {
        if (a.x != b.x) return a.x < b.x;
        if (a.y != b.y) return a.y < b.y;
        if (a.size != b.size) return a.size > b.size;
        if (a.angle != b.angle) return a.angle < b.angle;
        if (a.response != b.response) return a.response > b.response;
        return a.octave > b.octave;
    };
    bool radix_ok = true;
    for (int i = 0; i < nk && radix_ok; i++) radix_ok = kp[i].x >= 0.f && kp[i].y >= 0.f;      // also false for NaN
    if (radix_ok && nk > 1) {
        std::vector<uint64_t> key(nk), key2(nk);
        std::vector<int> id(nk), id2(nk);
        for (int i = 0; i < nk; i++) {
            uint32_t bx, by;
            const float fx = kp[i].x + 0.f, fy = kp[i].y + 0.f;          // -0 -> +0
            memcpy(&bx, &fx, 4); memcpy(&by, &fy, 4);
            key[i] = ((uint64_t)bx << 32) | by;
            id[i] = i;
        }
        std::vector<int> cnt(65536);
        for (int pass = 0; pass < 4; pass++) {
            const int sh = 16 * pass;
            std::fill(cnt.begin(), cnt.end(), 0);
            for (int i = 0; i < nk; i++) cnt[(key[i] >> sh) & 0xffff]++;
            int acc = 0;
            for (int d = 0; d < 65536; d++) { const int c = cnt[d]; cnt[d] = acc; acc += c; }
            for (int i = 0; i < nk; i++) { const int d = cnt[(key[i] >> sh) & 0xffff]++; key2[d] = key[i]; id2[d] = id[i]; }
            key.swap(key2); id.swap(id2);
        }
        std::vector<Kp> sorted(nk);
        for (int i = 0; i < nk; i++) sorted[i] = kp[id[i]];
        for (int i = 0; i < nk;) {
            int j = i + 1;
            while (j < nk && key[j] == key[i]) j++;
            if (j - i > 1) std::sort(sorted.begin() + i, sorted.begin() + j, less);
            i = j;
        }
        kp.swap(sorted);
    } else {
        std::sort(kp.begin(), kp.end(), less);
    }
    if (nk > 1) {
        int i = 0;
        for (int j = 1; j < nk; j++) {
            const Kp &a = kp[i], &b = kp[j];
            if (a.x != b.x || a.y != b.y || a.size != b.size || a.angle != b.angle) kp[++i] = kp[j];
        }
        kp.resize(i + 1);
    }
    if (nfeatures > 0 && (int)kp.size() > nfeatures) {
        std::vector<float> resp(kp.size());
        for (size_t i = 0; i < kp.size(); i++) resp[i] = kp[i].response;
        std::nth_element(resp.begin(), resp.begin() + (nfeatures - 1), resp.end(), std::greater<float>());
        const float thr = resp[nfeatures - 1];
        kp.erase(std::remove_if(kp.begin(), kp.end(), [thr](const Kp& a) { return !(a.response >= thr); }), kp.end());
    }
    *n = (int)kp.size();
    const int m = std::min((int)kp.size(), max_keypoints);
    for (int i = 0; i < m; i++) {
        xy[2 * i] = kp[i].x * 0.5f; xy[2 * i + 1] = kp[i].y * 0.5f;
        size[i] = kp[i].size * 0.5f;
        angle[i] = kp[i].angle;
        response[i] = kp[i].response;
        octave[i] = (kp[i].octave & ~255) | ((kp[i].octave - 1) & 255);
    }
    return FM3D_OK;
}

extern "C" {

int fm3d_detect_sift(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, int n_octave_layers,
                     double contrast_threshold, double edge_threshold, double sigma, int max_keypoints, float* xy, float* size,
                     float* angle, float* response, int32_t* octave, int* n) {
    return fm3d_detect_sift_impl(ctx, img, w, h, stride, nfeatures, n_octave_layers, contrast_threshold, edge_threshold, sigma, max_keypoints,
                                 xy, size, angle, response, octave, n, nullptr);
}

}  // extern "C"
