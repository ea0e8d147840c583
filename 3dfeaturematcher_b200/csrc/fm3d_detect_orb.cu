// fm3d_detect_orb.cu -- K15: cv::ORB's own detector and its descriptors on the pyramid levels (fm3d_detect_orb).
//
// Replaces feature_detector_->detect + descriptor_extractor_->compute of DescriptorsMatcher::compareWithNNDR / compare /
// crosscompare (DescriptorsMatcher/descriptorsmatcher.cpp:110-115, :91-96, :76-81) for DetectorType ORB + ExtractorType ORB
// (:273-279, :336-342: cv::ORB(OrbDetector.NumFeatures, ScaleFactor, NumLevels), everything else at cv::ORB's defaults).  The
// algorithm is OpenCV's (third party, unpinned in the reference; restated in oracle/orb_detect_np.py, which reproduces
// cv2.ORB_create(...).detectAndCompute keypoint for keypoint):
//   pyramid      level l = cv::resize(level l - 1, INTER_LINEAR_EXACT) to cvRound(w / s^l) x cvRound(h / s^l): 8.8 fixed-point
//                weights, rows then columns, one rounding -- orb_resize_kernel, bit-exact
//   per level    cv::FAST(20, nonmax) (K10, fm3d_detect_fast_dev), runByImageBorder(31), retainBest(2 n_l) by FAST score,
//                HarrisResponses (7 x 7 block of integer gradients, orb_harris_kernel), retainBest(n_l) by it,
//                ICAngles (integer moments of the radius-15 disc, orb_angle_kernel: one warp per keypoint)
//   descriptors  K13 (orb_blur_kernel + orb_kp_kernel) on the level image at the keypoint's level position
// The two retainBest selections run on the host between the kernels (a few thousand keypoints per level): they keep every
// keypoint that ties with the last kept response, as std::nth_element + std::partition do; OpenCV's order after them is
// unspecified, the output here is sorted by (level, y, x).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "fm3d_internal.cuh"

extern "C" int fm3d_detect_fast_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int threshold, int nonmax,
                                    int max_keypoints, float* xy, float* response, int* n_dev);
extern "C" int fm3d_describe_keypoints_orb_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, const float* kps, int n,
                                               uint8_t* descriptors, uint8_t* kept);

namespace {

constexpr int ORB_HALF_PATCH = 15;
constexpr int ORB_MAX_LEVELS = 16;

// INTER_LINEAR_EXACT for CV_8UC1: tap offsets and 8.8 weights of the second tap per destination column / row (host tables)
__global__ void orb_resize_kernel(const uint8_t* __restrict__ src, int sw, int sh, int sstride, const int* __restrict__ ox,
                                  const int* __restrict__ ax, const int* __restrict__ oy, const int* __restrict__ ay,
                                  uint8_t* __restrict__ dst, int dw, int dh) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const int x0 = ox[x], x1 = min(x0 + 1, sw - 1), a = ax[x];
    const int y0 = oy[y], y1 = min(y0 + 1, sh - 1), b = ay[y];
    const uint8_t* r0 = src + (size_t)y0 * sstride;
    const uint8_t* r1 = src + (size_t)y1 * sstride;
    const int h0 = r0[x0] * (256 - a) + r0[x1] * a;         // 8.8
    const int h1 = r1[x0] * (256 - a) + r1[x1] * a;
    const int v = h0 * (256 - b) + h1 * b;                   // 16.16
    dst[(size_t)y * dw + x] = (uint8_t)min(255, (v + (1 << 15)) >> 16);
}

// HarrisResponses (orb.cpp), blockSize 7, k = 0.04: integer sums of Ix^2, Iy^2, Ix Iy over the block, then float arithmetic
__global__ void orb_harris_kernel(const uint8_t* __restrict__ img, int w, int h, const int* __restrict__ xy, int n,
                                  float* __restrict__ resp) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int x0 = xy[2 * k], y0 = xy[2 * k + 1];
    int a = 0, b = 0, c = 0;
    for (int dy = -3; dy <= 3; dy++) {
        const uint8_t* p = img + (size_t)(y0 + dy) * w + x0;
        for (int dx = -3; dx <= 3; dx++) {
            const uint8_t* q = p + dx;
            const int Ix = ((int)q[1] - (int)q[-1]) * 2 + ((int)q[-w + 1] - (int)q[-w - 1]) + ((int)q[w + 1] - (int)q[w - 1]);
            const int Iy = ((int)q[w] - (int)q[-w]) * 2 + ((int)q[w - 1] - (int)q[-w - 1]) + ((int)q[w + 1] - (int)q[-w + 1]);
            a += Ix * Ix; b += Iy * Iy; c += Ix * Iy;
        }
    }
    const float scale = 1.f / ((1 << 2) * 7 * 255.f);
    const float scale_sq_sq = scale * scale * scale * scale;
    const float af = (float)a, bf = (float)b, cf = (float)c;
    resp[k] = __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(af, bf), __fmul_rn(cf, cf)), __fmul_rn(__fmul_rn(0.04f, __fadd_rn(af, bf)), __fadd_rn(af, bf))),
                        scale_sq_sq);
    (void)h;
}

__device__ __forceinline__ float orb_fast_atan2_deg(float y, float x) {    // cv::fastAtan2
    const float s = (float)(180.0 / M_PI);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s, p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, 2.220446049250313e-16f));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, 2.220446049250313e-16f));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

struct UMax { int u[ORB_HALF_PATCH + 2]; };

// ICAngles: m10 = sum u I, m01 = sum v I over the disc (rows v = -15 .. 15, half widths umax[|v|]); one warp per keypoint,
// lane = row (31 rows), integer warp sums
__global__ void orb_angle_kernel(const uint8_t* __restrict__ img, int w, const int* __restrict__ xy, int n, UMax um,
                                 float* __restrict__ angle) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= n) return;
    const int x0 = xy[2 * k], y0 = xy[2 * k + 1];
    int m10 = 0, m01 = 0;
    if (lane < 2 * ORB_HALF_PATCH + 1) {
        const int v = lane - ORB_HALF_PATCH;
        const int d = um.u[v < 0 ? -v : v];
        const uint8_t* p = img + (size_t)(y0 + v) * w + x0;
        int s = 0;
        for (int u = -d; u <= d; u++) { const int val = p[u]; m10 += u * val; s += val; }
        m01 = v * s;
    }
    for (int o = 16; o > 0; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    if (lane == 0) angle[k] = orb_fast_atan2_deg((float)m01, (float)m10);
}

// (x, y, 31, angle) rows for the descriptor stage from the level positions and the angles: KeyPoint::size of a level keypoint is
// the patch size, the angle is the one orb_angle_kernel left
__global__ void orb_k4_kernel(const int* __restrict__ xy, const float* __restrict__ angle, int n, float* __restrict__ k4) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    k4[4 * k] = (float)xy[2 * k]; k4[4 * k + 1] = (float)xy[2 * k + 1]; k4[4 * k + 2] = 31.f; k4[4 * k + 3] = angle[k];
}

inline int cv_round(double v) { return (int)std::nearbyint(v); }

void resize_tables(int ssize, int dsize, std::vector<int>& ofs, std::vector<int>& a1) {
    const double scale = (double)ssize / dsize;
    ofs.resize(dsize); a1.resize(dsize);
    for (int d = 0; d < dsize; d++) {
        double f = (d + 0.5) * scale - 0.5;
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0; }
        ofs[d] = s;
        a1[d] = cv_round(f * 256);
    }
}

// KeyPointsFilter::retainBest as a set: everything >= the n-th largest response
void retain_best(std::vector<int>& idx, const std::vector<float>& resp, int n) {
    if ((int)idx.size() <= n) return;
    if (n <= 0) { idx.clear(); return; }
    std::vector<float> r(idx.size());
    for (size_t i = 0; i < idx.size(); i++) r[i] = resp[idx[i]];
    std::nth_element(r.begin(), r.begin() + (n - 1), r.end(), std::greater<float>());
    const float thr = r[n - 1];
    idx.erase(std::remove_if(idx.begin(), idx.end(), [&](int i) { return !(resp[i] >= thr); }), idx.end());
}

}  // namespace

extern "C" {

int fm3d_detect_orb(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int nfeatures, double scale_factor, int nlevels,
                    int fast_threshold, int max_keypoints, float* xy, float* size, float* angle, float* response, int32_t* octave,
                    uint8_t* descriptors, int* n) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, img && n && w >= 2 && h >= 2 && stride >= w && nfeatures >= 0 && scale_factor > 1.0 && nlevels >= 1 &&
                            nlevels <= ORB_MAX_LEVELS && max_keypoints >= 0);
    FM3D_CHECK_ARG(ctx, max_keypoints == 0 || (xy && size && angle && response && octave));
    if (int rc = fm3d_bind(ctx)) return rc;
    const int edge = 31;
    // level sizes and scales: getScale = (float)pow((double)scaleFactor, level), Size(cvRound(w / scale), cvRound(h / scale))
    const float sf = (float)scale_factor;
    float sc[ORB_MAX_LEVELS];
    int lw[ORB_MAX_LEVELS], lh[ORB_MAX_LEVELS];
    size_t loff[ORB_MAX_LEVELS], total = 0;
    for (int l = 0; l < nlevels; l++) {
        sc[l] = (float)std::pow((double)sf, (double)l);
        lw[l] = l == 0 ? w : cv_round(w / sc[l]);
        lh[l] = l == 0 ? h : cv_round(h / sc[l]);
        if (lw[l] < 1 || lh[l] < 1) { nlevels = l; break; }
        loff[l] = total;
        total += ((size_t)lw[l] * lh[l] + 255) & ~(size_t)255;
    }
    // features per level (computeKeyPoints)
    int npl[ORB_MAX_LEVELS];
    {
        const float factor = 1.f / sf;
        float nd = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
        int sum = 0;
        for (int l = 0; l < nlevels - 1; l++) {
            npl[l] = cv_round(nd);
            sum += npl[l];
            nd *= factor;
        }
        npl[nlevels - 1] = std::max(nfeatures - sum, 0);
    }
    UMax um;
    {
        const int hp = ORB_HALF_PATCH;
        const int vmax = (int)std::floor(hp * std::sqrt(2.f) / 2 + 1), vmin = (int)std::ceil(hp * std::sqrt(2.f) / 2);
        for (int v = 0; v <= hp + 1; v++) um.u[v] = 0;
        for (int v = 0; v <= vmax; v++) um.u[v] = cv_round(std::sqrt((double)hp * hp - v * v));
        for (int v = hp, v0 = 0; v >= vmin; --v) {
            while (um.u[v0] == um.u[v0 + 1]) ++v0;
            um.u[v] = v0;
            ++v0;
        }
    }
    // device buffers: pyramid (slot 12), resize tables + lists of every level (slot 13).  The levels are worked on TOGETHER, stage
    // by stage, so that a frame costs four host round trips (corner counts, corners, Harris measures, angles + rows) instead of
    // six per level: the two retainBest selections stay on the host (they keep everything that ties with the last kept response,
    // as std::nth_element + std::partition do), everything between them runs level after level on the stream without a wait.
    uint8_t* pyr = nullptr;
    if (int rc = fm3d_scratch(ctx, 12, total, (void**)&pyr)) return rc;
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(pyr, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    int capl[ORB_MAX_LEVELS];                 // FAST corners a level has room for
    size_t coff[ORB_MAX_LEVELS], cap_total = 0, tab_ints = 0, toff[ORB_MAX_LEVELS];
    for (int l = 0; l < nlevels; l++) {
        capl[l] = std::max(4096, (lw[l] * lh[l]) / 16);
        coff[l] = cap_total;
        cap_total += (size_t)capl[l];
        toff[l] = tab_ints;
        if (l > 0) tab_ints += 2 * (size_t)(lw[l] + lh[l]);
    }
    const size_t b_tab = al(sizeof(int) * std::max<size_t>(tab_ints, 1)), b_xy = al(sizeof(float) * 2 * cap_total), b_r = al(sizeof(float) * cap_total);
    const size_t b_ixy = al(sizeof(int) * 2 * cap_total), b_k4 = al(sizeof(float) * 4 * cap_total), b_d = al((size_t)32 * cap_total), b_kept = al(cap_total);
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 13, b_tab + b_xy + 3 * b_r + b_ixy + b_k4 + b_d + b_kept + 256, (void**)&d)) return rc;
    int* d_tab = reinterpret_cast<int*>(d);
    float* d_xy = reinterpret_cast<float*>(d + b_tab);
    float* d_resp = reinterpret_cast<float*>(d + b_tab + b_xy);
    float* d_harris = reinterpret_cast<float*>(d + b_tab + b_xy + b_r);
    float* d_angle = reinterpret_cast<float*>(d + b_tab + b_xy + 2 * b_r);
    int* d_ixy = reinterpret_cast<int*>(d + b_tab + b_xy + 3 * b_r);
    float* d_k4 = reinterpret_cast<float*>(d + b_tab + b_xy + 3 * b_r + b_ixy);
    uint8_t* d_desc = reinterpret_cast<uint8_t*>(d + b_tab + b_xy + 3 * b_r + b_ixy + b_k4);
    uint8_t* d_kept = d_desc + b_d;
    int* d_n = reinterpret_cast<int*>(d_kept + b_kept);      // one corner count per level

    struct Kp { float x, y, size, angle, response; int octave, lx, ly; uint8_t desc[32]; };
    std::vector<Kp> out;

    // ---- stage A: the pyramid and the FAST corners of every level
    std::vector<int> tab(tab_ints), tx, ta, ty, tb;
    for (int l = 1; l < nlevels; l++) {
        resize_tables(lw[l - 1], lw[l], tx, ta);
        resize_tables(lh[l - 1], lh[l], ty, tb);
        int* t = tab.data() + toff[l];
        std::copy(tx.begin(), tx.end(), t); std::copy(ta.begin(), ta.end(), t + lw[l]);
        std::copy(ty.begin(), ty.end(), t + 2 * lw[l]); std::copy(tb.begin(), tb.end(), t + 2 * lw[l] + lh[l]);
    }
    if (tab_ints) if (int rc = fm3d_h2d(ctx, d_tab, tab.data(), sizeof(int) * tab_ints)) return rc;      // `tab` lives until the first wait below
    FM3D_CUDA(ctx, cudaMemsetAsync(d_n, 0, sizeof(int) * ORB_MAX_LEVELS, ctx->stream));
    ctx->n_copy++;
    bool active[ORB_MAX_LEVELS];
    for (int l = 0; l < nlevels; l++) {
        uint8_t* lim = pyr + loff[l];
        if (l > 0) {
            const int* t = d_tab + toff[l];
            const dim3 blk(32, 8);
            orb_resize_kernel<<<dim3((lw[l] + 31) / 32, (lh[l] + 7) / 8), blk, 0, ctx->stream>>>(
                pyr + loff[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], t, t + lw[l], t + 2 * lw[l], t + 2 * lw[l] + lh[l], lim, lw[l], lh[l]);
            FM3D_LAUNCH_CHECK(ctx);
        }
        active[l] = !(npl[l] == 0 || lw[l] <= 2 * edge || lh[l] <= 2 * edge);
        if (!active[l]) continue;
        if (int rc = fm3d_detect_fast_dev(ctx, lim, lw[l], lh[l], lw[l], fast_threshold, 1, capl[l], d_xy + 2 * coff[l], d_resp + coff[l], d_n + l)) return rc;
    }
    int nf[ORB_MAX_LEVELS] = {0};
    if (int rc = fm3d_d2h(ctx, nf, d_n, sizeof(int) * (size_t)nlevels)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::vector<float> hxy[ORB_MAX_LEVELS], hresp[ORB_MAX_LEVELS];
    for (int l = 0; l < nlevels; l++) {
        if (!active[l] || nf[l] == 0) continue;
        if (nf[l] > capl[l]) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "ORB: %d FAST corners on level %d exceed the buffer of %d", nf[l], l, capl[l]);
        hxy[l].resize(2 * (size_t)nf[l]); hresp[l].resize(nf[l]);
        if (int rc = fm3d_d2h(ctx, hxy[l].data(), d_xy + 2 * coff[l], sizeof(float) * 2 * (size_t)nf[l])) return rc;
        if (int rc = fm3d_d2h(ctx, hresp[l].data(), d_resp + coff[l], sizeof(float) * (size_t)nf[l])) return rc;
    }
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));

    // ---- stage B: runByImageBorder(edgeThreshold), retainBest(2 n_l) by FAST score; the Harris measure of the survivors
    std::vector<int> hixy;                    // survivors of all levels, level after level
    size_t moff[ORB_MAX_LEVELS + 1] = {0};
    for (int l = 0; l < nlevels; l++) {
        moff[l + 1] = moff[l];
        if (!active[l] || nf[l] == 0) continue;
        std::vector<int> idx;
        idx.reserve(nf[l]);
        for (int i = 0; i < nf[l]; i++) {
            const int x = (int)hxy[l][2 * i], y = (int)hxy[l][2 * i + 1];
            if (x >= edge && x < lw[l] - edge && y >= edge && y < lh[l] - edge) idx.push_back(i);
        }
        retain_best(idx, hresp[l], 2 * npl[l]);
        for (int i : idx) { hixy.push_back((int)hxy[l][2 * i]); hixy.push_back((int)hxy[l][2 * i + 1]); }
        moff[l + 1] = moff[l] + idx.size();
    }
    const size_t m_all = moff[nlevels];
    std::vector<float> hharris(m_all);
    if (m_all) {
        if (int rc = fm3d_h2d(ctx, d_ixy, hixy.data(), sizeof(int) * 2 * m_all)) return rc;
        for (int l = 0; l < nlevels; l++) {
            const int m = (int)(moff[l + 1] - moff[l]);
            if (m == 0) continue;
            orb_harris_kernel<<<(m + 127) / 128, 128, 0, ctx->stream>>>(pyr + loff[l], lw[l], lh[l], d_ixy + 2 * moff[l], m, d_harris + moff[l]);
            FM3D_LAUNCH_CHECK(ctx);
        }
        if (int rc = fm3d_d2h(ctx, hharris.data(), d_harris, sizeof(float) * m_all)) return rc;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }

    // ---- stage C: retainBest(n_l) by the Harris measure; angles and rows of what is left
    std::vector<int> kxy;                     // kept keypoints of all levels (level positions), level after level
    std::vector<float> kresp;
    size_t koff[ORB_MAX_LEVELS + 1] = {0};
    for (int l = 0; l < nlevels; l++) {
        koff[l + 1] = koff[l];
        const int m = (int)(moff[l + 1] - moff[l]);
        if (m == 0) continue;
        std::vector<float> hl(hharris.begin() + moff[l], hharris.begin() + moff[l + 1]);
        std::vector<int> idx2(m);
        for (int i = 0; i < m; i++) idx2[i] = i;
        retain_best(idx2, hl, npl[l]);
        for (int i : idx2) { kxy.push_back(hixy[2 * (moff[l] + i)]); kxy.push_back(hixy[2 * (moff[l] + i) + 1]); kresp.push_back(hl[i]); }
        koff[l + 1] = koff[l] + idx2.size();
    }
    const size_t k_all = koff[nlevels];
    std::vector<float> hangle(k_all);
    std::vector<uint8_t> hdesc((size_t)32 * k_all, 0);
    if (k_all) {
        if (int rc = fm3d_h2d(ctx, d_ixy, kxy.data(), sizeof(int) * 2 * k_all)) return rc;
        for (int l = 0; l < nlevels; l++) {
            const int m2 = (int)(koff[l + 1] - koff[l]);
            if (m2 == 0) continue;
            orb_angle_kernel<<<(m2 + 3) / 4, 128, 0, ctx->stream>>>(pyr + loff[l], lw[l], d_ixy + 2 * koff[l], m2, um, d_angle + koff[l]);
            FM3D_LAUNCH_CHECK(ctx);
        }
        if (descriptors) {
            orb_k4_kernel<<<(unsigned)((k_all + 255) / 256), 256, 0, ctx->stream>>>(d_ixy, d_angle, (int)k_all, d_k4);
            FM3D_LAUNCH_CHECK(ctx);
            for (int l = 0; l < nlevels; l++) {
                const int m2 = (int)(koff[l + 1] - koff[l]);
                if (m2 == 0) continue;
                if (int rc = fm3d_describe_keypoints_orb_dev(ctx, pyr + loff[l], lw[l], lh[l], lw[l], d_k4 + 4 * koff[l], m2, d_desc + 32 * koff[l],
                                                             d_kept + koff[l])) return rc;
            }
            if (int rc = fm3d_d2h(ctx, hdesc.data(), d_desc, (size_t)32 * k_all)) return rc;
        }
        if (int rc = fm3d_d2h(ctx, hangle.data(), d_angle, sizeof(float) * k_all)) return rc;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    out.reserve(k_all);
    for (int l = 0; l < nlevels; l++) {
        for (size_t i = koff[l]; i < koff[l + 1]; i++) {
            Kp k;
            k.lx = kxy[2 * i]; k.ly = kxy[2 * i + 1];
            k.x = (float)k.lx * sc[l]; k.y = (float)k.ly * sc[l];
            k.size = 31.f * sc[l];
            k.angle = hangle[i]; k.response = kresp[i]; k.octave = l;
            memcpy(k.desc, &hdesc[(size_t)32 * i], 32);
            out.push_back(k);
        }
    }
    std::sort(out.begin(), out.end(), [](const Kp& a, const Kp& b) {
        if (a.octave != b.octave) return a.octave < b.octave;
        if (a.ly != b.ly) return a.ly < b.ly;
        return a.lx < b.lx;
    });
    *n = (int)out.size();
    const int m = std::min((int)out.size(), max_keypoints);
    for (int i = 0; i < m; i++) {
        xy[2 * i] = out[i].x; xy[2 * i + 1] = out[i].y;
        size[i] = out[i].size; angle[i] = out[i].angle; response[i] = out[i].response; octave[i] = out[i].octave;
        if (descriptors) memcpy(descriptors + (size_t)32 * i, out[i].desc, 32);
    }
    return FM3D_OK;
}

}  // extern "C"
