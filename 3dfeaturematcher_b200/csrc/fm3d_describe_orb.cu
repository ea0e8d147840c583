// fm3d_describe_orb.cu -- K13: ORB (rBRIEF) descriptors at the keypoints of a whole frame.
//
// Replaces descriptor_extractor_->compute(frame, keypoints, descriptors) of
// DescriptorsMatcher::compareWithNNDR / compare / crosscompare
// (DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) for ExtractorType ORB (:336-342:
// cv::ORB(OrbDetector.NumFeatures, ScaleFactor, NumLevels): the knobs steer ORB's own detector and pyramid, which a
// provided keypoint of octave 0 -- what DetectorType FAST produces -- does not touch), the 32-byte binary rows the
// north star's Hamming sweep is quoted on.  OpenCV is a third-party dependency of the reference; the published
// algorithm (Rublee et al. 2011; modules/features2d/src/orb.cpp) is
//   * keypoints whose rounded position is within edgeThreshold = 31 pixels of the border are removed;
//   * the frame is blurred (7 x 7 Gaussian, sigma 2, reflect-101; observed through cv2's bits: the float convolution
//     rounded to u8);
//   * bit k = [B(c + R a_k) < B(c + R b_k)] for the 256 learned point pairs, R = rotation by KeyPoint::angle, every
//     rotated coordinate rounded to the nearest pixel.
// The pair table is not OpenCV source: tools/recover_orb_pattern.py recovers it from cv2.ORB_create().compute as a
// black box (exactly one pair per bit is consistent with 400 observations) and writes fm3d_orb_pattern.h.
//
// Kernels: orb_blur_kernel (both 7-tap passes of a 64 x 32 tile fused in shared memory, u8 -> u8), orb_kp_kernel (one
// warp per keypoint, one byte of the row per lane).
#include "fm3d_internal.cuh"

#include <math.h>

#include "fm3d_orb_pattern.h"

namespace {

constexpr int OB_KHALF = 3, OB_W = 64, OB_H = 32, OB_NT = 256, OB_EDGE = 31, OB_WARPS = 8;

struct OrbBlurArgs { float kern[OB_KHALF + 1]; };

__device__ signed char d_orb_pattern[256][4];

__device__ __forceinline__ int ob_reflect(int i, int n) {
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

__global__ void __launch_bounds__(OB_NT)
orb_blur_kernel(const OrbBlurArgs A, const uint8_t* __restrict__ img, int w, int h, int stride, uint8_t* __restrict__ out) {
    constexpr int IW = OB_W + 2 * OB_KHALF, IH = OB_H + 2 * OB_KHALF;
    __shared__ float in[IH][IW + 1];
    __shared__ float rowp[IH][OB_W];
    const int x0 = blockIdx.x * OB_W, y0 = blockIdx.y * OB_H, tid = threadIdx.x;
    for (int i = tid; i < IH * IW; i += OB_NT) {
        const int r = i / IW, c = i - r * IW;
        in[r][c] = fm3d_u8f(img[(size_t)ob_reflect(y0 + r - OB_KHALF, h) * stride + ob_reflect(x0 + c - OB_KHALF, w)]);
    }
    __syncthreads();
    for (int i = tid; i < IH * OB_W; i += OB_NT) {
        const int r = i / OB_W, c = i - r * OB_W;
        const float* px = &in[r][c + OB_KHALF];
        float s = __fmul_rn(A.kern[0], px[0]);
#pragma unroll
        for (int k = 1; k <= OB_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(px[-k], px[k]), s);
        rowp[r][c] = s;
    }
    __syncthreads();
    for (int i = tid; i < OB_H * OB_W; i += OB_NT) {
        const int r = i / OB_W, c = i - r * OB_W;
        const int gy = y0 + r, gx = x0 + c;
        if (gy >= h || gx >= w) continue;
        float s = __fmul_rn(A.kern[0], rowp[r + OB_KHALF][c]);
#pragma unroll
        for (int k = 1; k <= OB_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(rowp[r + OB_KHALF - k][c], rowp[r + OB_KHALF + k][c]), s);
        out[(size_t)gy * w + gx] = (uint8_t)fminf(fmaxf(rintf(s), 0.0f), 255.0f);        // saturate_cast<uchar>
    }
}

__global__ void __launch_bounds__(OB_WARPS * 32)
orb_kp_kernel(const uint8_t* __restrict__ blur, int w, int h, const float* __restrict__ kps, int n, uint8_t* __restrict__ desc,
              uint8_t* __restrict__ kept) {
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * OB_WARPS + (threadIdx.x >> 5);
    if (f >= n) return;
    const float x = kps[4 * (size_t)f], y = kps[4 * (size_t)f + 1], angle = kps[4 * (size_t)f + 3];
    // KeyPointsFilter::runByImageBorder: Rect(31, 31, w - 62, h - 62).contains(Point(pt)) -- the ROUNDED position
    const bool finite = isfinite(x) && isfinite(y) && isfinite(angle) && fabsf(x) < 1e8f && fabsf(y) < 1e8f;
    const int cx = finite ? __float2int_rn(x) : -1, cy = finite ? __float2int_rn(y) : -1;
    const bool keep = finite && cx >= OB_EDGE && cx < w - OB_EDGE && cy >= OB_EDGE && cy < h - OB_EDGE;
    if (!keep) {
        desc[(size_t)f * 32 + lane] = 0;
        if (lane == 0) kept[f] = 0;
        return;
    }
    const float rad = __fmul_rn(angle, (float)(M_PI / 180.0));
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);       // cosf / sinf: fp64 rounded once
    const uint8_t* center = blur + (size_t)cy * w + cx;
    unsigned byte = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int k = lane * 8 + i;
        const float ax = (float)d_orb_pattern[k][0], ay = (float)d_orb_pattern[k][1];
        const float bx = (float)d_orb_pattern[k][2], by = (float)d_orb_pattern[k][3];
        const int iax = __float2int_rn(__fsub_rn(__fmul_rn(ax, a), __fmul_rn(ay, b))), iay = __float2int_rn(__fadd_rn(__fmul_rn(ax, b), __fmul_rn(ay, a)));
        const int ibx = __float2int_rn(__fsub_rn(__fmul_rn(bx, a), __fmul_rn(by, b))), iby = __float2int_rn(__fadd_rn(__fmul_rn(bx, b), __fmul_rn(by, a)));
        const int t0 = center[iay * w + iax], t1 = center[iby * w + ibx];
        byte |= (t0 < t1 ? 1u : 0u) << i;
    }
    desc[(size_t)f * 32 + lane] = (uint8_t)byte;
    if (lane == 0) kept[f] = 1;
}

// keypoint k of extractDescriptorsFromPatches on the patches stacked into one S-wide image: (S/2, k S + S/2), size S, angle -1
__global__ void orb_patch_centres_kernel(float* __restrict__ kps, int n, int S) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float c = (float)(S / 2);
    kps[4 * (size_t)k] = c;
    kps[4 * (size_t)k + 1] = (float)k * (float)S + c;
    kps[4 * (size_t)k + 2] = (float)S;
    kps[4 * (size_t)k + 3] = -1.0f;
}

}  // namespace

extern "C" {

int fm3d_describe_keypoints_orb_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                    uint8_t* descriptors, uint8_t* kept) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && descriptors && kept && w >= 2 && h >= 2 && stride >= w)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    static_assert(sizeof(FM3D_ORB_PATTERN) == 1024, "256 pairs of two points");
    FM3D_CUDA(ctx, cudaMemcpyToSymbolAsync(d_orb_pattern, FM3D_ORB_PATTERN, sizeof(FM3D_ORB_PATTERN), 0, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    uint8_t* blur = nullptr;
    if (int rc = fm3d_scratch(ctx, 10, (size_t)w * h, (void**)&blur)) return rc;
    OrbBlurArgs A;
    {   // cv::getGaussianKernel(7, 2, CV_32F)
        double t[7], sum = 0;
        for (int i = 0; i < 7; i++) { const double x = i - 3; t[i] = exp(-0.5 / 4.0 * x * x); sum += t[i]; }
        for (int i = 0; i <= OB_KHALF; i++) A.kern[i] = (float)(t[OB_KHALF + i] / sum);
    }
    dim3 grid((w + OB_W - 1) / OB_W, (h + OB_H - 1) / OB_H);
    orb_blur_kernel<<<grid, OB_NT, 0, ctx->stream>>>(A, img, w, h, stride, blur);
    FM3D_LAUNCH_CHECK(ctx);
    orb_kp_kernel<<<(n + OB_WARPS - 1) / OB_WARPS, OB_WARPS * 32, 0, ctx->stream>>>(blur, w, h, kps, n, descriptors, kept);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_describe_patches_orb_dev(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, uint8_t* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && S >= 1 && (n == 0 || (patches && descriptors)));
    if (n == 0) return FM3D_OK;
    // cv::ORB keeps the keypoint (S/2, S/2) only if 31 <= S/2 < S - 31; the compared pixels (centre +- 13) and their 7 x 7
    // blur support then lie inside the patch, so the patches can be described as ONE image of n S rows: nothing a patch's
    // row depends on crosses into its neighbours, and the border rule on the stacked image is the per-patch rule.
    if (S / 2 < OB_EDGE || S / 2 >= S - OB_EDGE)
        return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "patch edge %d: cv::ORB removes the keypoint of extractDescriptorsFromPatches (needs 31 <= S/2 < S - 31)", S);
    FM3D_CHECK_ARG(ctx, (double)n * S < 2147483647.0 && (double)n * S < 16777216.0);      // row coordinates stay exact in float
    if (int rc = fm3d_bind(ctx)) return rc;
    char* d = nullptr;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    if (int rc = fm3d_scratch(ctx, 11, al(sizeof(float) * 4 * (size_t)n) + al((size_t)n), (void**)&d)) return rc;
    float* kps = reinterpret_cast<float*>(d);
    uint8_t* kept = reinterpret_cast<uint8_t*>(d + al(sizeof(float) * 4 * (size_t)n));
    orb_patch_centres_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(kps, n, S);
    FM3D_LAUNCH_CHECK(ctx);
    return fm3d_describe_keypoints_orb_dev(ctx, patches, S, n * S, S, kps, n, descriptors, kept);
}

int fm3d_describe_patches_orb(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, uint8_t* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && S >= 1 && (n == 0 || (patches && descriptors)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bp = (size_t)n * S * S, bd = 32 * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bp) + al(bd), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, patches, bp)) return rc;
    if (int rc = fm3d_describe_patches_orb_dev(ctx, reinterpret_cast<const uint8_t*>(d), n, S, reinterpret_cast<uint8_t*>(d + al(bp)))) return rc;
    if (int rc = fm3d_d2h(ctx, descriptors, d + al(bp), bd)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_describe_keypoints_orb(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                uint8_t* descriptors, uint8_t* kept) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && descriptors && kept && w >= 2 && h >= 2 && stride >= w)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bi = (size_t)w * h, bk = sizeof(float) * 4 * (size_t)n, bd = 32 * (size_t)n, bf = (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bi) + al(bk) + al(bd) + al(bf), (void**)&d)) return rc;
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(d, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    char* d_k = d + al(bi);
    char* d_d = d_k + al(bk);
    char* d_f = d_d + al(bd);
    if (int rc = fm3d_h2d(ctx, d_k, kps, bk)) return rc;
    if (int rc = fm3d_describe_keypoints_orb_dev(ctx, reinterpret_cast<const uint8_t*>(d), w, h, w, reinterpret_cast<const float*>(d_k), n,
                                                 reinterpret_cast<uint8_t*>(d_d), reinterpret_cast<uint8_t*>(d_f))) return rc;
    if (int rc = fm3d_d2h(ctx, descriptors, d_d, bd)) return rc;
    if (int rc = fm3d_d2h(ctx, kept, d_f, bf)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

}  // extern "C"
