// fm3d_describe_kp.cu -- K11: SIFT descriptors at the keypoints of a whole frame.
//
// Replaces descriptor_extractor_->compute(frame, keypoints, descriptors) of
// DescriptorsMatcher::compareWithNNDR / compare / crosscompare
// (DescriptorsMatcher/descriptorsmatcher.cpp:114-115, :95-96, :80-81) for ExtractorType SIFT (:302-314) and
// keypoints of octave 0 (what cv::FastFeatureDetector produces, :215-222: size 7, angle -1).  With such
// keypoints cv::SIFT::compute builds no scale space: every descriptor is read from
//     base = GaussianBlur(float(gray), sigma = sqrt(1.6^2 - 0.5^2))            (createInitialImage)
// by calcSIFTDescriptor(base, round(pt), ori = 360 - angle, scl = size / 2, d = 4, n = 8).
//
// Two kernels:
//   sift_base_kernel      u8 frame -> float base image, separable 13-tap blur fused in shared memory
//                         (HBM-bound: 1 B read + 4 B written per pixel).
//   describe_kp_kernel    four warps per keypoint, one WARP PER ROW OF CELLS of the 4 x 4 descriptor grid.
//                         A warp walks the bounding box of its row's trilinear support (rbin within one
//                         cell of the row), evaluates calcSIFTDescriptor's per-pixel arithmetic and adds the
//                         shares of its own row to per-lane PRIVATE accumulators in shared memory
//                         (4 cells x 8 orientations per lane, bank = lane: no conflicts, no atomics, fixed
//                         summation order => bit-reproducible).  A pixel is evaluated by the two warps
//                         whose rows it touches; the base image is L1/L2-resident.
// OpenCV addresses its (d+2)(d+2)(n+2) histogram flat.  A keypoint with angle -1 has ori = 361, so a
// gradient orientation below 1 degree keeps o0 = -1 after calcSIFTDescriptor's single wrap and its v0
// share lands in slot n+1 of the previous COLUMN cell, which the circular fold adds to that cell's
// orientation bin 1 (see fm3d_describe.cu).  A warp owns a whole row of cells, so that share stays inside the
// warp ("spill" below); the share of column 0 falls into the padding and is dropped, as in OpenCV.
#include <algorithm>
#include <climits>
#include <vector>

#include "fm3d_internal.cuh"

#include <math.h>

namespace {

constexpr int KP_KHALF = 6;          // ksize 13 = cvRound(8 sigma + 1) | 1 for sigma 1.5199
constexpr int BT_W = 64, BT_H = 32;  // base-image tile per CTA
constexpr int BT_NT = 256;
constexpr int KP_PER_CTA = 2;        // keypoints per CTA
constexpr int KP_NT = KP_PER_CTA * 128;   // four warps per keypoint: one per ROW of descriptor cells

struct BlurArgs {
    float kern[KP_KHALF + 1];        // kern[i] = coefficient at distance i from the centre tap
};

__device__ __forceinline__ int reflect101_any(int i, int n) {
    // cv::borderInterpolate(BORDER_REFLECT_101) for any offset (n >= 2)
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        if (i >= n) i = 2 * (n - 1) - i;
    }
    return i;
}

__global__ void __launch_bounds__(BT_NT)
sift_base_kernel(const BlurArgs A, const uint8_t* __restrict__ img, int w, int h, int stride, float* __restrict__ base) {
    constexpr int IW = BT_W + 2 * KP_KHALF, IH = BT_H + 2 * KP_KHALF;      // 76 x 44: rows of 304 B, 16-byte aligned
    __shared__ __align__(16) float in[IH][IW];
    __shared__ __align__(16) float rowp[IH][BT_W];
    const int x0 = blockIdx.x * BT_W, y0 = blockIdx.y * BT_H, tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    // interior tiles (no reflection, 4-byte aligned rows): 32-bit loads of columns x0 - 8 .. x0 + 71, of which x0 - 6 .. x0 + 69 are kept
    const bool interior = x0 >= 8 && x0 + BT_W + 8 <= w && y0 >= KP_KHALF && y0 + BT_H + KP_KHALF <= h &&
                          (stride & 3) == 0 && (reinterpret_cast<uintptr_t>(img) & 3) == 0;
    if (interior) {
        constexpr int Q = (BT_W + 16) / 4;       // 20 words per row
        for (int i = tid; i < IH * Q; i += BT_NT) {
            const int r = i / Q, q = i - r * Q;
            const unsigned v = *reinterpret_cast<const unsigned*>(img + (size_t)(y0 + r - KP_KHALF) * stride + (x0 - 8 + 4 * q));
            const int c = 4 * q - 2;             // column of the word's first byte in `in`
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (c + b >= 0 && c + b < IW) in[r][c + b] = fm3d_u8f((v >> (8 * b)) & 255u);
        }
    } else {
        for (int r = wid; r < IH; r += BT_NT / 32) {
            const uint8_t* src = img + (size_t)reflect101_any(y0 + r - KP_KHALF, h) * stride;
            for (int c = lane; c < IW; c += 32) in[r][c] = fm3d_u8f(src[reflect101_any(x0 + c - KP_KHALF, w)]);
        }
    }
    __syncthreads();
    // row pass, symmetric form k0 x0 + sum_i k_i (x_-i + x_+i) as cv::sepFilter2D's symmetric row filter.  Register-tiled like
    // K9's passes: four outputs per thread from four 16-byte loads instead of 13 loads per output (the kernel is bound by the
    // shared-memory pipe and instruction issue, not by HBM: profiles/r01i_sift_base_4k_kernel_ncu_*).
    for (int i = tid; i < IH * (BT_W / 4); i += BT_NT) {
        const int r = i / (BT_W / 4), c4 = i - r * (BT_W / 4);
        const float4* q = reinterpret_cast<const float4*>(&in[r][4 * c4]);
        float v[16];
#pragma unroll
        for (int k = 0; k < 4; k++) { const float4 t = q[k]; v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w; }
        float4 o;
        float* op = &o.x;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float s = __fmul_rn(A.kern[0], v[KP_KHALF + j]);
#pragma unroll
            for (int k = 1; k <= KP_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(v[KP_KHALF + j - k], v[KP_KHALF + j + k]), s);
            op[j] = s;
        }
        *reinterpret_cast<float4*>(&rowp[r][4 * c4]) = o;
    }
    __syncthreads();
    // column pass: four consecutive rows per thread from 16 loads
    for (int i = tid; i < (BT_H / 4) * BT_W; i += BT_NT) {
        const int r4 = i / BT_W, c = i - r4 * BT_W;
        const int gx = x0 + c;
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = rowp[4 * r4 + k][c];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int gy = y0 + 4 * r4 + j;
            float s = __fmul_rn(A.kern[0], v[KP_KHALF + j]);
#pragma unroll
            for (int k = 1; k <= KP_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(v[KP_KHALF + j - k], v[KP_KHALF + j + k]), s);
            if (gy < h && gx < w) base[(size_t)gy * w + gx] = s;
        }
    }
}

// cv::fastAtan2 (degrees in [0, 360)); same evaluation as K9 (fm3d_describe.cu)
__device__ __forceinline__ float kp_fast_atan2_deg(float y, float x) {
    const float s = 57.29577951308232f;
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    const bool ge = ax >= ay;
    const float c = __fdividef(ge ? ay : ax, __fadd_rn(ge ? ax : ay, eps)), c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    if (!ge) a = __fsub_rn(90.0f, a);
    if (x < 0) a = __fsub_rn(180.0f, a);
    if (y < 0) a = __fsub_rn(360.0f, a);
    return a;
}

__device__ __forceinline__ float kp_sqrt_approx(float a) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

struct KpGeom {
    float cos_t, sin_t;      // divided by hist_width
    float ori;
    int ptx, pty, radius;
    int valid;
    const float* img;        // the Gaussian image the descriptor is read from, and its size
    int w, h;
};

// Keypoints that carry an octave (cv::SIFT's own detector): the image of the keypoint's octave / layer in the Gaussian
// pyramid (float offset from its start), its size, and the factor from frame coordinates to that image
// (calcDescriptors: unpackOctave, ptf = kpt.pt * scale, size = kpt.size * scale).
struct KpLayer {
    int off, w, h;
    float scale;
};

__global__ void __launch_bounds__(KP_NT)
describe_kp_kernel(const float* __restrict__ base, int fw, int fh, const float* __restrict__ kps, const KpLayer* __restrict__ layers,
                   int n, float* __restrict__ desc) {
    __shared__ KpGeom G[KP_PER_CTA];
    __shared__ float acc[KP_NT / 32][32][32];        // [warp][cell column * 8 + orientation bin][lane]: bank = lane
    __shared__ float hist[KP_PER_CTA][128];
    __shared__ float scale_thr[KP_PER_CTA][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = warp >> 2, ci = warp & 3;       // keypoint of this CTA, row of descriptor cells
    const size_t f = (size_t)blockIdx.x * KP_PER_CTA + slot;
    if ((tid & 127) == 0) {
        KpGeom g;
        g.valid = 0;
        if (f < (size_t)n) {
            float x = kps[4 * f], y = kps[4 * f + 1], size = kps[4 * f + 2];
            const float angle = kps[4 * f + 3];
            // DescriptorExtractor::compute drops keypoints outside the image or of size <= FLT_EPSILON before SIFT
            // sees them (KeyPointsFilter::runByImageBorder / runByKeypointSize); such rows are written as zeros
            g.valid = (x >= 0.f && y >= 0.f && x < (float)fw && y < (float)fh && size > 1.1920929e-07f && isfinite(size) && isfinite(angle)) ? 1 : 0;
            g.img = base; g.w = fw; g.h = fh;
            if (layers) {
                const KpLayer L = layers[f];
                g.img = base + L.off; g.w = L.w; g.h = L.h;
                x = __fmul_rn(x, L.scale); y = __fmul_rn(y, L.scale); size = __fmul_rn(size, L.scale);
            }
            const int w = g.w, h = g.h;
            g.ptx = __float2int_rn(x); g.pty = __float2int_rn(y);                  // cvRound
            float ori = __fsub_rn(360.0f, angle);
            if (fabsf(ori - 360.0f) < 1.1920929e-07f) ori = 0.0f;
            const float scl = __fmul_rn(size, 0.5f);
            const float rad = __fmul_rn(ori, (float)(M_PI / 180));
            // cosf / sinf of the C library are correctly rounded for almost every argument: fp64 evaluation rounded once
            const float cos_t = (float)cos((double)rad), sin_t = (float)sin((double)rad);
            const float hist_width = __fmul_rn(3.0f, scl);
            const int radius = __float2int_rn(__fmul_rn(__fmul_rn(__fmul_rn(hist_width, 1.4142135623730951f), 5.0f), 0.5f));
            const int diag = (int)sqrt((double)w * w + (double)h * h);
            g.radius = radius < diag ? radius : diag;
            g.cos_t = __fdiv_rn(cos_t, hist_width);
            g.sin_t = __fdiv_rn(sin_t, hist_width);
            g.ori = ori;
        }
        G[slot] = g;
    }
    float* my = &acc[warp][0][lane];                 // bin b of this lane: my[b * 32]
#pragma unroll
    for (int b = 0; b < 32; b++) my[b * 32] = 0.0f;
    __syncthreads();
    const KpGeom g = G[slot];
    const float* __restrict__ img = g.img;
    const int w = g.w, h = g.h;
    if (g.valid) {
        // bounding box (pixel offsets i = row, j = column) of the band rbin in [ci-1, ci+1), cbin in (-1, 4);
        // *_rot = *bin - 1.5
        int i_lo, i_hi, j_lo, j_hi;
        {
            const float inv = 1.0f / (g.cos_t * g.cos_t + g.sin_t * g.sin_t);
            const float rr[2] = {(float)ci - 2.5f, (float)ci - 0.5f};
            const float cc[2] = {-2.5f, 2.5f};
            float fi_lo = 1e30f, fi_hi = -1e30f, fj_lo = 1e30f, fj_hi = -1e30f;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const float fj = (cc[b] * g.cos_t + rr[a] * g.sin_t) * inv;
                    const float fi = (-cc[b] * g.sin_t + rr[a] * g.cos_t) * inv;
                    fi_lo = fminf(fi_lo, fi); fi_hi = fmaxf(fi_hi, fi);
                    fj_lo = fminf(fj_lo, fj); fj_hi = fmaxf(fj_hi, fj);
                }
            const float R = (float)g.radius;
            i_lo = (int)floorf(fmaxf(fi_lo, -R - 1.0f)) - 1; i_hi = (int)ceilf(fminf(fi_hi, R + 1.0f)) + 1;
            j_lo = (int)floorf(fmaxf(fj_lo, -R - 1.0f)) - 1; j_hi = (int)ceilf(fminf(fj_hi, R + 1.0f)) + 1;
            i_lo = max(i_lo, max(-g.radius, 1 - g.pty)); i_hi = min(i_hi, min(g.radius, h - 2 - g.pty));
            j_lo = max(j_lo, max(-g.radius, 1 - g.ptx)); j_hi = min(j_hi, min(g.radius, w - 2 - g.ptx));
        }
        const int bw = j_hi - j_lo + 1, bh = i_hi - i_lo + 1;
        if (bw > 0 && bh > 0) {
            const long long total = (long long)bw * bh;
            int q = lane / bw, rem = lane - q * bw;
            const int step_q = 32 / bw, step_r = 32 - step_q * bw;
            for (long long k = lane; k < total; k += 32, q += step_q, rem += step_r) {
                if (rem >= bw) { rem -= bw; q++; }
                const int i = i_lo + q, j = j_lo + rem;
                const float fi = (float)i, fj = (float)j;
                const float c_rot = __fsub_rn(__fmul_rn(fj, g.cos_t), __fmul_rn(fi, g.sin_t));
                const float r_rot = __fadd_rn(__fmul_rn(fj, g.sin_t), __fmul_rn(fi, g.cos_t));
                const float rbin = __fsub_rn(__fadd_rn(r_rot, 2.0f), 0.5f);
                const float cbin = __fsub_rn(__fadd_rn(c_rot, 2.0f), 0.5f);
                if (!(rbin > -1.0f && rbin < 4.0f && cbin > -1.0f && cbin < 4.0f)) continue;
                const float r0f = floorf(rbin), c0f = floorf(cbin);
                const int dr = ci - (int)r0f;            // this row of cells is row r0 (dr = 0) or r0 + 1 (dr = 1) of the split
                if (dr != 0 && dr != 1) continue;
                const int c0 = (int)c0f;                 // -1 .. 3
                const float* p = img + (size_t)(g.pty + i) * w + (g.ptx + j);
                const float dx = __fsub_rn(__ldg(p + 1), __ldg(p - 1)), dy = __fsub_rn(__ldg(p - w), __ldg(p + w));
                const float wgt = __expf(__fmul_rn(__fadd_rn(__fmul_rn(c_rot, c_rot), __fmul_rn(r_rot, r_rot)), -0.125f));
                const float mag = __fmul_rn(kp_sqrt_approx(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))), wgt);
                float obin = __fmul_rn(__fsub_rn(kp_fast_atan2_deg(dy, dx), g.ori), 8 / 360.0f);
                const float of = floorf(obin);
                obin = __fsub_rn(obin, of);
                int o0 = (int)of;
                o0 = o0 < 0 ? o0 + 8 : o0;
                o0 = o0 >= 8 ? o0 - 8 : o0;
                const float rb = __fsub_rn(rbin, r0f), cb = __fsub_rn(cbin, c0f);
                const float v_r1 = __fmul_rn(mag, rb);
                const float v_r = dr ? v_r1 : __fsub_rn(mag, v_r1);
                const float v_rc1 = __fmul_rn(v_r, cb), v_rc0 = __fsub_rn(v_r, v_rc1);
                if (o0 >= 0) {
                    // slot o0 and slot o0 + 1 (slot n folds onto bin 0) of the cells in columns c0 and c0 + 1
                    const int oa = (o0 & 7) * 32, ob = ((o0 + 1) & 7) * 32;
                    if (c0 >= 0) {
                        const float v1 = __fmul_rn(v_rc0, obin), v0 = __fsub_rn(v_rc0, v1);
                        float* cell = my + c0 * 256;
                        cell[oa] = __fadd_rn(cell[oa], v0);
                        cell[ob] = __fadd_rn(cell[ob], v1);
                    }
                    if (c0 < 3) {
                        const float v1 = __fmul_rn(v_rc1, obin), v0 = __fsub_rn(v_rc1, v1);
                        float* cell = my + (c0 + 1) * 256;
                        cell[oa] = __fadd_rn(cell[oa], v0);
                        cell[ob] = __fadd_rn(cell[ob], v1);
                    }
                } else {
                    // o0 == -1 (ori = 361, orientation below 1 degree): the v1 share is slot 0 of the proper cell, the v0
                    // share is flat slot n+1 of the cell one column to the LEFT, which the circular fold adds to its bin 1
#pragma unroll
                    for (int dc = 0; dc < 2; dc++) {
                        const int col = c0 + dc;
                        const float v_rc = dc ? v_rc1 : v_rc0;
                        const float v1 = __fmul_rn(v_rc, obin), v0 = __fsub_rn(v_rc, v1);
                        if (col >= 0 && col <= 3) my[col * 256] = __fadd_rn(my[col * 256], v1);
                        if (col >= 1 && col <= 4) my[(col - 1) * 256 + 32] = __fadd_rn(my[(col - 1) * 256 + 32], v0);
                    }
                }
            }
        }
    }
    __syncwarp();
    {   // lane L sums bin L over the 32 private copies, skewed so that every lane reads another bank; fixed order
        const float* row = &acc[warp][lane][0];
        float s = 0.0f;
#pragma unroll
        for (int k = 0; k < 32; k++) s = __fadd_rn(s, row[(k + lane) & 31]);
        hist[slot][ci * 32 + lane] = s;              // (ci * 4 + column) * 8 + orientation
    }
    __syncthreads();
    // normalisation (calcSIFTDescriptor's tail): clip at 0.2 |h|, scale to 512, saturate to u8
    if (ci == 0) {
        float v[4], s = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; k++) { v[k] = hist[slot][lane * 4 + k]; s = fmaf(v[k], v[k], s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float thr = __fmul_rn(__fsqrt_rn(s), 0.2f);
        s = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; k++) { const float c = fminf(v[k], thr); s = fmaf(c, c, s); }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            scale_thr[slot][0] = __fdiv_rn(512.0f, fmaxf(__fsqrt_rn(s), 1.1920929e-07f));
            scale_thr[slot][1] = thr;
        }
    }
    __syncthreads();
    if (f < (size_t)n) {
        const int e = tid & 127;
        desc[f * 128 + e] = g.valid ? fminf(fmaxf(rintf(__fmul_rn(fminf(hist[slot][e], scale_thr[slot][1]), scale_thr[slot][0])), 0.0f), 255.0f) : 0.0f;
    }
}

void blur_args(BlurArgs& A) {
    const double sigma = sqrt(fmax(1.6 * 1.6 - 0.5 * 0.5, 0.01));       // createInitialImage, no up-sampling
    double t[2 * KP_KHALF + 1], sum = 0;
    for (int i = 0; i <= 2 * KP_KHALF; i++) {                            // cv::getGaussianKernel(13, sigma, CV_32F)
        const double x = i - KP_KHALF;
        t[i] = exp(-0.5 / (sigma * sigma) * x * x);
        sum += t[i];
    }
    for (int i = 0; i <= KP_KHALF; i++) A.kern[i] = (float)(t[KP_KHALF + i] / sum);
}

}  // namespace

extern "C" {

int fm3d_sift_base_image_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, float* base) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, img && base && w >= 2 && h >= 2 && stride >= w);
    if (int rc = fm3d_bind(ctx)) return rc;
    BlurArgs A;
    blur_args(A);
    dim3 grid((w + BT_W - 1) / BT_W, (h + BT_H - 1) / BT_H);
    sift_base_kernel<<<grid, BT_NT, 0, ctx->stream>>>(A, img, w, h, stride, base);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_describe_keypoints_sift_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                     float* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && descriptors && w >= 2 && h >= 2 && stride >= w)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    float* base = nullptr;
    if (int rc = fm3d_scratch(ctx, 8, sizeof(float) * (size_t)w * h, (void**)&base)) return rc;
    if (int rc = fm3d_sift_base_image_dev(ctx, img, w, h, stride, base)) return rc;
    describe_kp_kernel<<<(n + KP_PER_CTA - 1) / KP_PER_CTA, KP_NT, 0, ctx->stream>>>(base, w, h, kps, nullptr, n, descriptors);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_describe_keypoints_sift(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                 float* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && descriptors && w >= 2 && h >= 2 && stride >= w)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bi = (size_t)w * h, bk = sizeof(float) * 4 * (size_t)n, bd = sizeof(float) * 128 * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bi) + al(bk) + al(bd), (void**)&d)) return rc;
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(d, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    if (int rc = fm3d_h2d(ctx, d + al(bi), kps, bk)) return rc;
    float* d_desc = reinterpret_cast<float*>(d + al(bi) + al(bk));
    if (int rc = fm3d_describe_keypoints_sift_dev(ctx, reinterpret_cast<const uint8_t*>(d), w, h, w,
                                                  reinterpret_cast<const float*>(d + al(bi)), n, d_desc)) return rc;
    if (int rc = fm3d_d2h(ctx, descriptors, d_desc, bd)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

// cv::SIFT::compute on keypoints that carry an octave, on a pyramid that is already on the device (first octave P.first_octave):
// every descriptor from the layer (octave - first octave)(nOctaveLayers + 3) + layer of its keypoint.  Scratch 0 holds the
// keypoints, their layers and the descriptors (the pyramid lives in scratch 12).
static int describe_on_pyramid(fm3d_ctx* ctx, const fm3d_sift_pyramid& P, int w, int h, const float* kps, const int32_t* octaves, int n,
                               int n_octave_layers, float* descriptors) {
    const int first = P.first_octave;
    std::vector<KpLayer> layers(n);
    for (int i = 0; i < n; i++) {
        int o = octaves[i] & 255;
        const int l = (octaves[i] >> 8) & 255;
        o = o < 128 ? o : (-128 | o);
        if (l > n_octave_layers + 2) return fm3d_fail(ctx, FM3D_ERR_INVALID_ARG, "SIFT keypoint %d: layer %d of %d", i, l, n_octave_layers + 3);
        const int oi = o - first;
        if (oi < 0 || oi >= P.n_octaves) return fm3d_fail(ctx, FM3D_ERR_INVALID_ARG, "SIFT keypoint %d: octave %d beyond the pyramid of this frame", i, o);
        const size_t off = P.gauss_off[oi] + (size_t)l * P.w[oi] * P.h[oi];
        if (off > (size_t)INT_MAX) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "SIFT pyramid larger than 2^31 floats");
        layers[i] = KpLayer{(int)off, P.w[oi], P.h[oi], o >= 0 ? 1.f / (float)(1 << o) : (float)(1 << -o)};
    }
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bk = sizeof(float) * 4 * (size_t)n, bl = sizeof(KpLayer) * (size_t)n, bd = sizeof(float) * 128 * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bk) + al(bl) + al(bd), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, kps, bk)) return rc;
    if (int rc = fm3d_h2d(ctx, d + al(bk), layers.data(), bl)) return rc;
    float* d_desc = reinterpret_cast<float*>(d + al(bk) + al(bl));
    describe_kp_kernel<<<(n + KP_PER_CTA - 1) / KP_PER_CTA, KP_NT, 0, ctx->stream>>>(
        P.gauss, w, h, reinterpret_cast<const float*>(d), reinterpret_cast<const KpLayer*>(d + al(bk)), n, d_desc);
    FM3D_LAUNCH_CHECK(ctx);
    if (int rc = fm3d_d2h(ctx, descriptors, d_desc, bd)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));     // `layers` must outlive the copy
    return FM3D_OK;
}

int fm3d_describe_keypoints_sift_oct(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps,
                                     const int32_t* octaves, int n, int n_octave_layers, double sigma, float* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (img && kps && octaves && descriptors && w >= 2 && h >= 2 && stride >= w)));
    FM3D_CHECK_ARG(ctx, n_octave_layers >= 1 && sigma > 0);
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    // detectAndCompute(useProvidedKeypoints): the octave range of the keypoints decides the pyramid
    int first = 0, last = INT_MIN;
    for (int i = 0; i < n; i++) {
        int o = octaves[i] & 255;
        o = o < 128 ? o : (-128 | o);
        first = std::min(first, o); last = std::max(last, o);
    }
    if (first < -1) return fm3d_fail(ctx, FM3D_ERR_INVALID_ARG, "SIFT keypoints: first octave %d < -1", first);
    uint8_t* d_img = nullptr;
    if (int rc = fm3d_scratch(ctx, 15, (size_t)w * h, (void**)&d_img)) return rc;
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(d_img, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    fm3d_sift_pyramid P;
    if (int rc = fm3d_sift_build_pyramid(ctx, d_img, w, h, w, first, last - first + 1, n_octave_layers, sigma, false, &P)) return rc;
    return describe_on_pyramid(ctx, P, w, h, kps, octaves, n, n_octave_layers, descriptors);
}

int fm3d_detect_and_describe_sift(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, int n_octave_layers,
                                  double contrast_threshold, double edge_threshold, double sigma, int max_keypoints, float* xy, float* size,
                                  float* angle, float* response, int32_t* octave, int* n, float* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, max_keypoints == 0 || descriptors);
    fm3d_sift_pyramid P;
    if (int rc = fm3d_detect_sift_impl(ctx, img, w, h, stride, nfeatures, n_octave_layers, contrast_threshold, edge_threshold, sigma, max_keypoints,
                                       xy, size, angle, response, octave, n, &P)) return rc;
    const int m = std::min(*n, max_keypoints);
    if (m <= 0 || *n > max_keypoints) return FM3D_OK;        // too little room: the caller comes back with the real count
    // the keypoints in the layout of fm3d_describe_keypoints_sift_oct, described on the pyramid they were found on (the same
    // images cv::SIFT::compute would build again: first octave -1, every octave)
    std::vector<float> k4((size_t)4 * m);
    for (int i = 0; i < m; i++) { k4[4 * i] = xy[2 * i]; k4[4 * i + 1] = xy[2 * i + 1]; k4[4 * i + 2] = size[i]; k4[4 * i + 3] = angle[i]; }
    return describe_on_pyramid(ctx, P, w, h, k4.data(), octave, m, n_octave_layers, descriptors);
}

}  // extern "C"
