// fm3d_describe_brisk.cu -- K12: BRISK descriptors at the keypoints of a whole frame.
//
// Replaces descriptor_extractor_->compute(frame, keypoints, descriptors) of
// DescriptorsMatcher::compareWithNNDR / compare / crosscompare
// (DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) for ExtractorType BRISK (:343-349:
// cv::BRISK(BriskDetector.Threshold, BriskDetector.Octaves): both knobs only steer BRISK's own detector, which the
// reference never runs), the binary extractor that feeds the Hamming matcher (:64-67).  OpenCV is a third-party
// dependency of the reference; the published algorithm (modules/features2d/src/brisk.cpp; Leutenegger et al. 2011) is
//   * a 60-point pattern on 5 rings, 64 scales x 1024 rotations, box smoothing of half width sigma per ring,
//     512 short pairs (the bits) and 870 long pairs (the orientation);
//   * per keypoint: scale index from KeyPoint::size; keypoints nearer to the border than the pattern extent are
//     removed; 60 smoothed intensities (integer box integration with sub-pixel borders, over the image for small
//     boxes and over the integral image for large ones) at rotation 0 -> orientation; at the rotated pattern -> bits.
// All integer work except the pattern coordinates: bits identical to cv2.BRISK_create().compute
// (tests/golden/brisk_keypoints.npz).
//
// Kernels: integral image (row scan + column scan, int32 as cv::integral's CV_32S), then ONE WARP PER KEYPOINT: every
// lane smooths two pattern points, the long-pair sums are integer warp reductions (order-independent, exact), the 512
// comparisons leave the warp as 16 ballots.  The pattern tables cv::BRISK precomputes (47 MB) are not stored: a
// point is scale * radius * cos(alpha + theta) evaluated in fp64 where it is needed; the per-scale constants come
// from the host (same libm as OpenCV's).
#include "fm3d_internal.cuh"

#include <math.h>

#include <mutex>
#include <vector>

namespace {

constexpr int BK_POINTS = 60, BK_SCALES = 64, BK_NROT = 1024, BK_SHORT = 512, BK_LONG_MAX = 1024;
constexpr int BK_WARPS = 4;

struct BriskTables {
    float scale[BK_SCALES];          // scaleList_
    float radius[BK_SCALES][5];      // scaleList_[s] * radiusList[ring] (float product, as generateKernel forms it)
    float sigma[BK_SCALES][5];       // half width of the smoothing box per ring
    int size[BK_SCALES];             // sizeList_: pattern extent = border
    int n_long;
};

struct BriskPairs {                  // device copy: short pairs (i, j) and long pairs (i, j, weighted dx, dy)
    uint8_t short_i[BK_SHORT], short_j[BK_SHORT];
    uint8_t long_i[BK_LONG_MAX], long_j[BK_LONG_MAX];
    int16_t long_wx[BK_LONG_MAX], long_wy[BK_LONG_MAX];
};

__constant__ int c_ring_n[5] = {1, 10, 14, 15, 20};
__constant__ int c_ring_first[5] = {0, 1, 11, 25, 40};

// ---------------------------------------------------------------- integral image (cv::integral, CV_32S)
__global__ void integral_rows_kernel(const uint8_t* __restrict__ img, int w, int h, int stride, int* __restrict__ integ) {
    // one warp per image row: integ[(y + 1) * (w + 1) + x + 1] = prefix sum of the row; row 0 and column 0 are zero
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int W1 = w + 1;
    if (y == 0) for (int x = lane; x < W1; x += 32) integ[x] = 0;
    if (y >= h) return;
    int carry = 0;
    int* out = integ + (size_t)(y + 1) * W1;
    if (lane == 0) out[0] = 0;
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int x = x0 + lane;
        int v = x < w ? (int)img[(size_t)y * stride + x] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (x < w) out[x + 1] = v + carry;
        carry += __shfl_sync(0xffffffffu, v, 31);
    }
}

__global__ void integral_cols_kernel(int w, int h, int* __restrict__ integ) {
    // one thread per column, rows in batches of 16: the loads of a batch are issued together (a load-add-store loop over
    // one array serialises on the L2 latency of every row: 0.18 ms for a 720p frame, measured), then the running sum
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int W1 = w + 1;
    if (x > w) return;
    int s = 0;
    int* col = integ + x;
    int y = 1;
    for (; y + 15 <= h; y += 16) {
        int v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = col[(size_t)(y + k) * W1];
#pragma unroll
        for (int k = 0; k < 16; k++) { s += v[k]; col[(size_t)(y + k) * W1] = s; }
    }
    for (; y <= h; y++) { s += col[(size_t)y * W1]; col[(size_t)y * W1] = s; }
}

// ---------------------------------------------------------------- smoothedIntensity (brisk.cpp), integer arithmetic
__device__ __forceinline__ int brisk_smoothed(const uint8_t* __restrict__ img, int w, int stride, const int* __restrict__ integ,
                                              float key_x, float key_y, float px, float py, float sigma_half) {
    const float xf = __fadd_rn(px, key_x), yf = __fadd_rn(py, key_y);
    const float area = __fmul_rn(__fmul_rn(4.0f, sigma_half), sigma_half);
    if (sigma_half < 0.5f) {
        const int x = (int)xf, y = (int)yf;
        const int r_x = (int)__fmul_rn(__fsub_rn(xf, (float)x), 1024.0f), r_y = (int)__fmul_rn(__fsub_rn(yf, (float)y), 1024.0f);
        const int r_x_1 = 1024 - r_x, r_y_1 = 1024 - r_y;
        const uint8_t* p = img + (size_t)y * stride + x;
        const int v = r_x_1 * r_y_1 * (int)p[0] + r_x * r_y_1 * (int)p[1] + r_x * r_y * (int)p[stride + 1] + r_x_1 * r_y * (int)p[stride];
        return (v + 512) / 1024;
    }
    const int scaling = (int)(4194304.0 / (double)area);
    const int scaling2 = (int)((double)__fmul_rn((float)scaling, area) / 1024.0);
    const float x_1 = __fsub_rn(xf, sigma_half), x1 = __fadd_rn(xf, sigma_half);
    const float y_1 = __fsub_rn(yf, sigma_half), y1 = __fadd_rn(yf, sigma_half);
    const int x_left = (int)((double)x_1 + 0.5), y_top = (int)((double)y_1 + 0.5);
    const int x_right = (int)((double)x1 + 0.5), y_bottom = (int)((double)y1 + 0.5);
    const float r_x_1 = __fadd_rn(__fsub_rn((float)x_left, x_1), 0.5f), r_y_1 = __fadd_rn(__fsub_rn((float)y_top, y_1), 0.5f);
    const float r_x1 = __fadd_rn(__fsub_rn(x1, (float)x_right), 0.5f), r_y1 = __fadd_rn(__fsub_rn(y1, (float)y_bottom), 0.5f);
    const int dx = x_right - x_left - 1, dy = y_bottom - y_top - 1;
    const float fs = (float)scaling;
    const int A = (int)__fmul_rn(__fmul_rn(r_x_1, r_y_1), fs), B = (int)__fmul_rn(__fmul_rn(r_x1, r_y_1), fs);
    const int C = (int)__fmul_rn(__fmul_rn(r_x1, r_y1), fs), D = (int)__fmul_rn(__fmul_rn(r_x_1, r_y1), fs);
    const int r_x_1_i = (int)__fmul_rn(r_x_1, fs), r_y_1_i = (int)__fmul_rn(r_y_1, fs);
    const int r_x1_i = (int)__fmul_rn(r_x1, fs), r_y1_i = (int)__fmul_rn(r_y1, fs);
    auto I = [&](int yy, int xx) { return (int)img[(size_t)yy * stride + xx]; };
    // two's-complement wrap-around like the reference's int arithmetic: accumulate in unsigned
    unsigned ret = (unsigned)A * I(y_top, x_left) + (unsigned)B * I(y_top, x_right) + (unsigned)C * I(y_bottom, x_right) +
                   (unsigned)D * I(y_bottom, x_left);
    if (dx + dy > 2) {
        const int W1 = w + 1;
        // sum of the image over rows ya .. yb-1, columns xa .. xb-1
        auto S = [&](int ya, int yb, int xa, int xb) {
            return integ[(size_t)yb * W1 + xb] - integ[(size_t)ya * W1 + xb] - integ[(size_t)yb * W1 + xa] + integ[(size_t)ya * W1 + xa];
        };
        ret += (unsigned)S(y_top, y_top + 1, x_left + 1, x_right) * (unsigned)r_y_1_i;
        ret += (unsigned)S(y_top + 1, y_bottom, x_left + 1, x_right) * (unsigned)scaling;
        ret += (unsigned)S(y_top + 1, y_bottom, x_left, x_left + 1) * (unsigned)r_x_1_i;
        ret += (unsigned)S(y_top + 1, y_bottom, x_right, x_right + 1) * (unsigned)r_x1_i;
        ret += (unsigned)S(y_bottom, y_bottom + 1, x_left + 1, x_right) * (unsigned)r_y1_i;
    } else {
        for (int xx = x_left + 1; xx < x_right; xx++) ret += (unsigned)r_y_1_i * I(y_top, xx);
        for (int yy = y_top + 1; yy < y_bottom; yy++) {
            ret += (unsigned)r_x_1_i * I(yy, x_left);
            for (int xx = x_left + 1; xx < x_right; xx++) ret += (unsigned)scaling * I(yy, xx);
            ret += (unsigned)r_x1_i * I(yy, x_right);
        }
        for (int xx = x_left + 1; xx < x_right; xx++) ret += (unsigned)r_y1_i * I(y_bottom, xx);
    }
    ret += (unsigned)(scaling2 / 2);
    return (int)ret / scaling2;
}

__device__ __forceinline__ void brisk_point(const BriskTables& T, int scale, int rot, int point, float& px, float& py, float& sg) {
    int ring = 0;
#pragma unroll
    for (int r = 1; r < 5; r++) if (point >= c_ring_first[r]) ring = r;
    const int num = point - c_ring_first[ring];
    const double alpha = (double)num * 2 * M_PI / (double)c_ring_n[ring];
    const double theta = (double)rot * 2 * M_PI / (double)BK_NROT;
    const double rad = (double)T.radius[scale][ring];
    px = (float)(rad * cos(alpha + theta));
    py = (float)(rad * sin(alpha + theta));
    sg = T.sigma[scale][ring];
}

__global__ void __launch_bounds__(BK_WARPS * 32)
brisk_kp_kernel(const BriskTables T, const BriskPairs* __restrict__ P, const uint8_t* __restrict__ img, int w, int h, int stride,
                const int* __restrict__ integ, const float* __restrict__ kps, int n, int compute_orientation,
                uint8_t* __restrict__ desc, uint8_t* __restrict__ kept, float* __restrict__ angles) {
    __shared__ int vals[BK_WARPS][64];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = blockIdx.x * BK_WARPS + wid;
    if (f >= n) return;                                   // whole warps leave: only __syncwarp below
    const float x = kps[4 * (size_t)f], y = kps[4 * (size_t)f + 1], size = kps[4 * (size_t)f + 2];
    float angle = kps[4 * (size_t)f + 3];
    // scale index (computeDescriptorsAndOrOrientation): float arithmetic, + 0.5 in double, truncated, saturated
    int scale;
    {
        const float log2f_ = 0.693147180559945f;
        const float lb_scalerange = __fdiv_rn(3.4011974334716797f, log2f_);       // logf(30.f) / log2, float division
        const float basic06 = __fmul_rn(12.0f, 0.6f);
        // std::log(float) is logf; evaluated in fp64 and rounded once (logf of the C library is correctly rounded)
        const float lg = (float)log((double)__fdiv_rn(size, basic06));
        const double v = (double)__fmul_rn(__fdiv_rn((float)BK_SCALES, lb_scalerange), __fdiv_rn(lg, log2f_)) + 0.5;
        scale = v > 0.0 ? (v < 1e9 ? (int)v : BK_SCALES) : 0;       // NaN / negative -> 0 as max((int)v, 0)
        if (scale >= BK_SCALES) scale = BK_SCALES - 1;
    }
    const int border = T.size[scale];
    // RoiPredicate(border, border, cols - border, rows - border): removed when outside [min, max)
    const bool keep = x >= (float)border && x < (float)(w - border) && y >= (float)border && y < (float)(h - border) && size > 0.0f;
    uint32_t* row = reinterpret_cast<uint32_t*>(desc + (size_t)f * 64);
    if (!keep) {
        if (lane < 16) row[lane] = 0u;
        if (lane == 0) { kept[f] = 0; angles[f] = angle; }
        return;
    }
    int* v = vals[wid];
    if (compute_orientation) {
        for (int p = lane; p < BK_POINTS; p += 32) {
            float px, py, sg;
            brisk_point(T, scale, 0, p, px, py, sg);
            v[p] = brisk_smoothed(img, w, stride, integ, x, y, px, py, sg);
        }
        __syncwarp();
        int d0 = 0, d1 = 0;
        for (int k = lane; k < T.n_long; k += 32) {
            const int dt = v[P->long_i[k]] - v[P->long_j[k]];
            d0 += dt * (int)P->long_wx[k] / 1024;         // C division: truncation toward zero
            d1 += dt * (int)P->long_wy[k] / 1024;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { d0 += __shfl_xor_sync(0xffffffffu, d0, o); d1 += __shfl_xor_sync(0xffffffffu, d1, o); }
        // the sums pass through float, atan2 and the conversion to degrees run in fp64, rounded once (pinned against cv2)
        angle = (float)(atan2((double)(float)d1, (double)(float)d0) / M_PI * 180.0);
        __syncwarp();
    }
    int theta;
    if (angle == -1.0f) theta = 0;
    else {
        theta = (int)((double)BK_NROT * ((double)angle / 360.0) + 0.5);
        if (theta < 0) theta += BK_NROT;
        if (theta >= BK_NROT) theta -= BK_NROT;
    }
    if (angle < 0.0f) angle = __fadd_rn(angle, 360.0f);
    for (int p = lane; p < BK_POINTS; p += 32) {
        float px, py, sg;
        brisk_point(T, scale, theta, p, px, py, sg);
        v[p] = brisk_smoothed(img, w, stride, integ, x, y, px, py, sg);
    }
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
        const int b = k * 32 + lane;                      // bit b of the row: word k, bit `lane`
        const unsigned word = __ballot_sync(0xffffffffu, v[P->short_i[b]] > v[P->short_j[b]]);
        if (lane == k) row[k] = word;
    }
    if (lane == 0) { kept[f] = 1; angles[f] = angle; }
}

// ---------------------------------------------------------------- host: the constants of generateKernel
struct HostBrisk {
    BriskTables T;
    BriskPairs P;
    bool ok = false;
};

const HostBrisk& host_brisk() {
    static HostBrisk H;
    static std::once_flag once;
    std::call_once(once, [] {
        const float rList[5] = {(float)(0.85 * 0.0), (float)(0.85 * 2.9), (float)(0.85 * 4.9), (float)(0.85 * 7.4), (float)(0.85 * 10.8)};
        const int nList[5] = {1, 10, 14, 15, 20};
        // scalerange_ is a float member of cv::BRISK, so std::log(scalerange_) is logf (pinned through the angles cv2 returns)
        const float lb_scale = (float)((double)logf(30.0f) / log(2.0));
        const float lb_scale_step = lb_scale / (float)BK_SCALES;
        const float sigma_scale = 1.3f;
        std::vector<float> x0(BK_POINTS), y0(BK_POINTS);
        for (int s = 0; s < BK_SCALES; s++) {
            const float sc = (float)pow(2.0, (double)((float)s * lb_scale_step));
            H.T.scale[s] = sc;
            H.T.size[s] = 0;
            int p = 0;
            for (int ring = 0; ring < 5; ring++) {
                const float rad = sc * rList[ring];
                H.T.radius[s][ring] = rad;
                H.T.sigma[s][ring] = ring == 0 ? sigma_scale * sc * 0.5f
                                               : (float)(sigma_scale * sc * (double)rList[ring] * sin(M_PI / nList[ring]));
                const int size = (int)ceil((double)(rad + H.T.sigma[s][ring])) + 1;
                if (H.T.size[s] < size) H.T.size[s] = size;
                for (int num = 0; num < nList[ring]; num++, p++) {
                    if (s != 0) continue;
                    const double alpha = (double)num * 2 * M_PI / (double)nList[ring];
                    x0[p] = (float)((double)rad * cos(alpha));
                    y0[p] = (float)((double)rad * sin(alpha));
                }
            }
        }
        const float dMax = 5.85f, dMin = 8.2f, dMin_sq = dMin * dMin, dMax_sq = dMax * dMax;
        int ns = 0, nl = 0;
        bool fits = true;
        for (int i = 1; i < BK_POINTS; i++)
            for (int j = 0; j < i; j++) {
                const float dx = x0[j] - x0[i], dy = y0[j] - y0[i];
                const float norm_sq = dx * dx + dy * dy;
                if (norm_sq > dMin_sq) {
                    if (nl < BK_LONG_MAX) {
                        H.P.long_i[nl] = (uint8_t)i; H.P.long_j[nl] = (uint8_t)j;
                        H.P.long_wx[nl] = (int16_t)(int)((dx / norm_sq) * 2048.0 + 0.5);
                        H.P.long_wy[nl] = (int16_t)(int)((dy / norm_sq) * 2048.0 + 0.5);
                    } else fits = false;
                    nl++;
                } else if (norm_sq < dMax_sq) {
                    if (ns < BK_SHORT) { H.P.short_i[ns] = (uint8_t)i; H.P.short_j[ns] = (uint8_t)j; } else fits = false;
                    ns++;
                }
            }
        H.T.n_long = nl;
        H.ok = fits && ns == BK_SHORT;      // cv::BRISK's default pattern: 512 short pairs = 64 bytes, 870 long pairs
    });
    return H;
}

}  // namespace

extern "C" {

int fm3d_describe_keypoints_brisk_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                      int compute_orientation, uint8_t* descriptors, uint8_t* kept, float* angles) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && descriptors && kept && angles && w >= 1 && h >= 1 && stride >= w)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    FM3D_CHECK_ARG(ctx, (double)w * h * 255.0 < 2147483648.0);     // the integral image is int32 like cv::integral's CV_32S
    const HostBrisk& H = host_brisk();
    if (!H.ok) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "BRISK pattern does not have 512 short pairs on this host's libm");
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char* d = nullptr;
    const size_t bi = sizeof(int) * (size_t)(w + 1) * (h + 1);
    if (int rc = fm3d_scratch(ctx, 9, al(sizeof(BriskPairs)) + al(bi), (void**)&d)) return rc;
    BriskPairs* d_pairs = reinterpret_cast<BriskPairs*>(d);
    int* integ = reinterpret_cast<int*>(d + al(sizeof(BriskPairs)));
    if (int rc = fm3d_h2d(ctx, d_pairs, &H.P, sizeof(BriskPairs))) return rc;   // 5 KB, static host source
    integral_rows_kernel<<<(h + 7) / 8, 256, 0, ctx->stream>>>(img, w, h, stride, integ);
    FM3D_LAUNCH_CHECK(ctx);
    integral_cols_kernel<<<(w + 1 + 31) / 32, 32, 0, ctx->stream>>>(w, h, integ);
    FM3D_LAUNCH_CHECK(ctx);
    brisk_kp_kernel<<<(n + BK_WARPS - 1) / BK_WARPS, BK_WARPS * 32, 0, ctx->stream>>>(H.T, d_pairs, img, w, h, stride, integ, kps, n,
                                                                                    compute_orientation ? 1 : 0, descriptors, kept, angles);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_describe_keypoints_brisk(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                  int compute_orientation, uint8_t* descriptors, uint8_t* kept, float* angles) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && descriptors && kept && angles && w >= 1 && h >= 1 && stride >= w)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bi = (size_t)w * h, bk = sizeof(float) * 4 * (size_t)n, bd = 64 * (size_t)n, bf = (size_t)n, ba = sizeof(float) * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bi) + al(bk) + al(bd) + al(bf) + al(ba), (void**)&d)) return rc;
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(d, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    char* d_k = d + al(bi);
    char* d_d = d_k + al(bk);
    char* d_f = d_d + al(bd);
    char* d_a = d_f + al(bf);
    if (int rc = fm3d_h2d(ctx, d_k, kps, bk)) return rc;
    if (int rc = fm3d_describe_keypoints_brisk_dev(ctx, reinterpret_cast<const uint8_t*>(d), w, h, w, reinterpret_cast<const float*>(d_k), n,
                                                   compute_orientation, reinterpret_cast<uint8_t*>(d_d), reinterpret_cast<uint8_t*>(d_f),
                                                   reinterpret_cast<float*>(d_a))) return rc;
    if (int rc = fm3d_d2h(ctx, descriptors, d_d, bd)) return rc;
    if (int rc = fm3d_d2h(ctx, kept, d_f, bf)) return rc;
    if (int rc = fm3d_d2h(ctx, angles, d_a, ba)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

}  // extern "C"
