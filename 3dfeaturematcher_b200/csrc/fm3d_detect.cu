// fm3d_detect.cu -- K10: FAST-9-16 corner detection (the stage right before the matcher).
//
// Replaces feature_detector_->detect(image, keypoints) of DescriptorsMatcher::compareWithNNDR
// (DescriptorsMatcher/descriptorsmatcher.cpp:110-111) for DetectorType FAST, DetectorMode STATIC (:215-222:
// cv::FastFeatureDetector(FeatureOptions.FastDetector.Threshold, NonMaxSuppression > 0)).  OpenCV is a third-party
// dependency of the reference; the published algorithm (modules/features2d/src/fast.cpp, fast_score.cpp) is:
//   * a pixel at least 3 pixels from the border is a corner if 9 contiguous pixels of the 16-pixel Bresenham circle of
//     radius 3 are all darker than v - t or all brighter than v + t (strict);
//   * with non-maximum suppression its score is cornerScore<16>: the largest threshold for which it is still a corner,
//     = max(t, max over the 16 arcs of min(v - x), max over the arcs of min(x - v)) - 1; it is kept iff that score is
//     strictly greater than the scores of its 8 neighbours (non-corners score 0); response = score.  Without
//     suppression every corner is reported with response 0;
//   * keypoints come out in row-major order: KeyPoint(x, y, size 7, angle -1, response).
// Integer work, bit-exact against cv2.FastFeatureDetector (tests/golden/fast_keypoints.npz).
//
// Three small kernels: scores (one pixel per thread, 32 x 8 tiles with a 3-pixel halo staged in shared memory: every
// image byte is read from HBM once, one score byte written per pixel), suppression + per-row counts + keep masks (one
// bit per pixel), and an order-preserving emit (one warp per row walks the set bits of the row's mask words behind the
// exclusive scan of the row counts).
#include "fm3d_internal.cuh"

namespace {

constexpr int FT_W = 32, FT_H = 8, FT_R = 3;

__constant__ int8_t c_circle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                        {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// smap: score + 1 for corners (<= 255), 0 elsewhere
__global__ void __launch_bounds__(FT_W * FT_H)
fast_score_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, int threshold, uint8_t* __restrict__ smap) {
    __shared__ uint8_t tile[FT_H + 2 * FT_R][FT_W + 2 * FT_R + 2];
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
    const int tx = threadIdx.x & (FT_W - 1), ty = threadIdx.x / FT_W;
    for (int i = threadIdx.x; i < (FT_H + 2 * FT_R) * (FT_W + 2 * FT_R); i += FT_W * FT_H) {
        const int r = i / (FT_W + 2 * FT_R), c = i - r * (FT_W + 2 * FT_R);
        const int gx = x0 + c - FT_R, gy = y0 + r - FT_R;
        tile[r][c] = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? img[(size_t)gy * pitch + gx] : 0;
    }
    __syncthreads();
    const int x = x0 + tx, y = y0 + ty;
    if (x >= w || y >= h) return;
    int out = 0;
    if (x >= FT_R && x < w - FT_R && y >= FT_R && y < h - FT_R) {
        const int v = tile[ty + FT_R][tx + FT_R];
        int d[16];
        unsigned brighter = 0u, darker = 0u;      // circle pixel brighter than v + t / darker than v - t
#pragma unroll
        for (int k = 0; k < 16; k++) {
            d[k] = v - (int)tile[ty + FT_R + c_circle[k][1]][tx + FT_R + c_circle[k][0]];
            darker |= (d[k] > threshold) ? (1u << k) : 0u;
            brighter |= (d[k] < -threshold) ? (1u << k) : 0u;
        }
        // 9 contiguous set bits on the 16-bit ring: AND of the mask with its 8 rotations
        auto arc9 = [](unsigned m) {
            m |= m << 16;
            unsigned a = m & (m >> 1);
            a &= a >> 2;           // 4 in a row
            a &= a >> 4;           // 8 in a row
            a &= m >> 8;           // 9 in a row
            return (a & 0xffffu) != 0u;
        };
        if (arc9(darker) || arc9(brighter)) {
            // cornerScore<16>: max over arcs of the minimum of d (darker arcs) and of -d (brighter arcs)
            int amax = -1000, bmin = 1000;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int mn = d[k], mx = d[k];
#pragma unroll
                for (int j = 1; j < 9; j++) { const int e = d[(k + j) & 15]; mn = min(mn, e); mx = max(mx, e); }
                amax = max(amax, mn);
                bmin = min(bmin, mx);
            }
            const int score = max(max(threshold, amax), -bmin) - 1;
            out = score + 1;
        }
    }
    smap[(size_t)y * w + x] = (uint8_t)out;
}

__device__ __forceinline__ bool fast_keep(const uint8_t* __restrict__ smap, int w, int h, int x, int y, int nonmax) {
    const int sp = smap[(size_t)y * w + x];
    if (sp == 0) return false;
    if (!nonmax) return true;
    const int s = sp - 1;
#pragma unroll
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
        for (int dx = -1; dx <= 1; dx++) {
            if (dx == 0 && dy == 0) continue;
            const int xx = x + dx, yy = y + dy;
            int n = 0;
            if (xx >= 0 && xx < w && yy >= 0 && yy < h) { const int q = smap[(size_t)yy * w + xx]; n = q > 0 ? q - 1 : 0; }
            if (!(s > n)) return false;
        }
    return true;
}

// one warp per row: number of keypoints of the row, and the keep decisions as one 32-bit mask per 32 pixels (the emit pass
// reads the masks instead of repeating the 9-pixel suppression test)
__global__ void fast_count_kernel(const uint8_t* __restrict__ smap, int w, int h, int nonmax, int* __restrict__ row_count,
                                  unsigned* __restrict__ mask) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= h) return;
    const int mw = (w + 31) >> 5;
    int cnt = 0;
    for (int x0 = 0; x0 < w; x0 += 32) {
        const int x = x0 + lane;
        const unsigned bal = __ballot_sync(0xffffffffu, x < w && fast_keep(smap, w, h, x, y, nonmax));
        if (lane == 0) mask[(size_t)y * mw + (x0 >> 5)] = bal;
        cnt += __popc(bal);
    }
    if (lane == 0) row_count[y] = cnt;
}

// exclusive scan of the row counts (one block), total into *n_out
__global__ void __launch_bounds__(1024) fast_scan_kernel(const int* __restrict__ row_count, int h, int* __restrict__ row_off, int* __restrict__ n_out) {
    __shared__ int part[1024];
    const int per = (h + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(h, lo + per);
    int s = 0;
    for (int y = lo; y < hi; y++) s += row_count[y];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < 1024; k++) { const int t = part[k]; part[k] = acc; acc += t; }
        *n_out = acc;
    }
    __syncthreads();
    int off = part[threadIdx.x];
    for (int y = lo; y < hi; y++) { row_off[y] = off; off += row_count[y]; }
}

// one warp per row: keypoints in ascending x behind the row's offset
__global__ void fast_emit_kernel(const uint8_t* __restrict__ smap, const unsigned* __restrict__ mask, int w, int h, int nonmax,
                                 const int* __restrict__ row_off, int max_kp, float* __restrict__ xy, float* __restrict__ response) {
    const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (y >= h) return;
    const int mw = (w + 31) >> 5;
    int base = row_off[y];
    for (int m0 = 0; m0 < mw; m0 += 32) {
        // 32 mask words per step; every lane then walks the set bits of the words in order
        const unsigned mine = m0 + lane < mw ? mask[(size_t)y * mw + m0 + lane] : 0u;
        int cnt = __popc(mine), off = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
        int o = base + off - cnt;                                   // exclusive prefix: first output slot of this lane's word
        unsigned bits = mine;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const int x = ((m0 + lane) << 5) + b;
            if (o < max_kp) {
                xy[2 * o] = (float)x; xy[2 * o + 1] = (float)y;
                response[o] = nonmax ? (float)((int)smap[(size_t)y * w + x] - 1) : 0.0f;
            }
            o++;
        }
        base += __shfl_sync(0xffffffffu, off, 31);
    }
}

}  // namespace

extern "C" {

int fm3d_detect_fast_dev(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int threshold, int nonmax,
                         int max_keypoints, float* xy, float* response, int* n_dev) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, img && w > 0 && h > 0 && stride >= w && max_keypoints >= 0 && n_dev && (max_keypoints == 0 || (xy && response)));
    if (int rc = fm3d_bind(ctx)) return rc;
    threshold = threshold < 0 ? 0 : (threshold > 255 ? 255 : threshold);      // as cv::FAST clamps it
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    char* d = nullptr;
    const size_t bm = sizeof(unsigned) * (size_t)((w + 31) >> 5) * h;
    if (int rc = fm3d_scratch(ctx, 7, al((size_t)w * h) + 2 * al(sizeof(int) * (size_t)h) + al(bm), (void**)&d)) return rc;
    uint8_t* smap = reinterpret_cast<uint8_t*>(d);
    int* row_count = reinterpret_cast<int*>(d + al((size_t)w * h));
    int* row_off = reinterpret_cast<int*>(d + al((size_t)w * h) + al(sizeof(int) * (size_t)h));
    unsigned* mask = reinterpret_cast<unsigned*>(d + al((size_t)w * h) + 2 * al(sizeof(int) * (size_t)h));
    dim3 grid((w + FT_W - 1) / FT_W, (h + FT_H - 1) / FT_H);
    fast_score_kernel<<<grid, FT_W * FT_H, 0, ctx->stream>>>(img, w, h, stride, threshold, smap);
    FM3D_LAUNCH_CHECK(ctx);
    fast_count_kernel<<<(h + 7) / 8, 256, 0, ctx->stream>>>(smap, w, h, nonmax ? 1 : 0, row_count, mask);
    FM3D_LAUNCH_CHECK(ctx);
    fast_scan_kernel<<<1, 1024, 0, ctx->stream>>>(row_count, h, row_off, n_dev);
    FM3D_LAUNCH_CHECK(ctx);
    if (max_keypoints > 0) {
        fast_emit_kernel<<<(h + 7) / 8, 256, 0, ctx->stream>>>(smap, mask, w, h, nonmax ? 1 : 0, row_off, max_keypoints, xy, response);
        FM3D_LAUNCH_CHECK(ctx);
    }
    return FM3D_OK;
}

int fm3d_detect_fast(fm3d_ctx* ctx, const uint8_t* img, int w, int h, int stride, int threshold, int nonmax,
                     int max_keypoints, float* xy, float* response, int* n) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, img && w > 0 && h > 0 && stride >= w && max_keypoints >= 0 && n && (max_keypoints == 0 || (xy && response)));
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bi = (size_t)w * h, bx = sizeof(float) * 2 * (size_t)max_keypoints, br = sizeof(float) * (size_t)max_keypoints;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bi) + al(bx) + al(br) + 256, (void**)&d)) return rc;
    // rows are packed on the way in (stride may exceed w)
    FM3D_CUDA(ctx, cudaMemcpy2DAsync(d, (size_t)w, img, (size_t)stride, (size_t)w, (size_t)h, cudaMemcpyHostToDevice, ctx->stream));
    ctx->n_copy++;
    float* d_xy = reinterpret_cast<float*>(d + al(bi));
    float* d_r = reinterpret_cast<float*>(d + al(bi) + al(bx));
    int* d_n = reinterpret_cast<int*>(d + al(bi) + al(bx) + al(br));
    if (int rc = fm3d_detect_fast_dev(ctx, reinterpret_cast<const uint8_t*>(d), w, h, w, threshold, nonmax, max_keypoints, d_xy, d_r, d_n)) return rc;
    int total = 0;
    if (int rc = fm3d_d2h(ctx, &total, d_n, sizeof(int))) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int got = total < max_keypoints ? total : max_keypoints;
    if (got > 0) {
        if (int rc = fm3d_d2h(ctx, xy, d_xy, sizeof(float) * 2 * (size_t)got)) return rc;
        if (int rc = fm3d_d2h(ctx, response, d_r, sizeof(float) * (size_t)got)) return rc;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    *n = total;      // > max_keypoints: the first max_keypoints (row-major order) were written
    return FM3D_OK;
}

}  // extern "C"
