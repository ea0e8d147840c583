// fm3d_pyramid.cu -- K4: image upload and the cv::pyrDown chain of NormalOptimizer::setImages /
// compute_pyramids (Triangulator/normaloptimizer.cpp:191-221).
//
// cv::pyrDown on CV_8U is integer arithmetic: separable [1 4 6 4 1] in both directions,
// BORDER_REFLECT_101, dst = (sum + 128) >> 8, dst size ((w+1)/2, (h+1)/2); this kernel is
// bit-exact with it.  HBM-bound: 1.25 bytes per source pixel per level.  One CTA produces a
// 64x32 output tile from a (2*64+8)x(2*32+3) source tile staged in shared memory with 32-bit
// loads; the horizontal pass goes to a u16 smem plane, the vertical pass packs four outputs
// into one 32-bit store.  Both images are processed by one launch per level (blockIdx.z).
//
// Every level is stored with a row pitch that is a multiple of 16 bytes so that the normal
// optimiser can address it with a TMA tensor map.
#include "fm3d_internal.cuh"

namespace {

constexpr int TW = 64, TH = 32;        // output tile
constexpr int SW = 2 * TW + 8;         // source tile width incl. alignment slop (136)
constexpr int SH = 2 * TH + 3;         // source tile height (67)
constexpr int SP = SW + 4;             // smem pitch of the source tile

__device__ __forceinline__ int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256)
pyrdown_kernel(const uint8_t* __restrict__ src1, const uint8_t* __restrict__ src2, int sw, int sh,
               int spitch, uint8_t* __restrict__ dst1, uint8_t* __restrict__ dst2, int dw, int dh,
               int dpitch) {
    __shared__ __align__(16) uint8_t s_src[SH * SP];
    __shared__ __align__(16) uint16_t s_h[SH * TW];
    const uint8_t* __restrict__ src = blockIdx.z ? src2 : src1;
    uint8_t* __restrict__ dst = blockIdx.z ? dst2 : dst1;
    const int ox0 = blockIdx.x * TW, oy0 = blockIdx.y * TH;
    const int sx0 = 2 * ox0 - 4, sy0 = 2 * oy0 - 2;  // sx0 is a multiple of 4
    const int tid = threadIdx.x;
    const bool interior = sx0 >= 0 && sy0 >= 0 && sx0 + SW <= sw && sy0 + SH <= sh;
    if (interior) {
        constexpr int WPR = SW / 4;  // 34 words per row
        for (int i = tid; i < SH * WPR; i += 256) {
            const int r = i / WPR, c = i - r * WPR;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(src + (size_t)(sy0 + r) * spitch + sx0 + 4 * c);
            *reinterpret_cast<uint32_t*>(&s_src[r * SP + 4 * c]) = v;
        }
    } else {
        for (int i = tid; i < SH * SW; i += 256) {
            const int r = i / SW, c = i - r * SW;
            const int yy = reflect101(sy0 + r, sh), xx = reflect101(sx0 + c, sw);
            s_src[r * SP + c] = src[(size_t)yy * spitch + xx];
        }
    }
    __syncthreads();
    // horizontal pass: output column x reads source columns 2x-2..2x+2 = tile columns 2x+2..2x+6
    for (int i = tid; i < SH * TW; i += 256) {
        const int r = i / TW, x = i - r * TW;
        const uint8_t* p = &s_src[r * SP + 2 * x + 2];
        s_h[r * TW + x] = (uint16_t)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();
    // vertical pass, four adjacent outputs per thread
    for (int i = tid; i < TH * (TW / 4); i += 256) {
        const int y = i / (TW / 4), xq = (i - y * (TW / 4)) * 4;
        const int oy = oy0 + y, ox = ox0 + xq;
        if (oy >= dh || ox >= dw) continue;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint16_t* p = &s_h[(2 * y) * TW + xq + k];
            const int s = p[0] + 4 * p[TW] + 6 * p[2 * TW] + 4 * p[3 * TW] + p[4 * TW];
            packed |= (uint32_t)((s + 128) >> 8) << (8 * k);
        }
        uint8_t* o = dst + (size_t)oy * dpitch + ox;
        if (ox + 4 <= dpitch) {
            *reinterpret_cast<uint32_t*>(o) = packed;  // pitch is a multiple of 16: the pad bytes are ours
        } else {
            for (int k = 0; k < 4 && ox + k < dw; k++) o[k] = (uint8_t)(packed >> (8 * k));
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Fused pyramid: ONE launch copies both frames into level 0 (when they come from caller-owned device memory) and
// builds up to three cv::pyrDown levels of each.  A CTA owns a B x B tile of the base level (B = 64) and everything above
// it: it stages the base pixels the tile's share of the top level depends on (85 x 85 for three levels), then computes
// level after level in shared memory, halos included (recomputed by the neighbouring CTAs: 1.76x the base reads, from L2),
// and writes only the pixels it owns.  Integer arithmetic, sum of the 25 taps before the single (s + 128) >> 8, so the
// result is bit-identical with the separable two-pass form (and with cv::pyrDown).  720p, 3 levels, both frames:
// 480 CTAs, one launch instead of three launches and two device-to-device copies.
constexpr int FB = 64;                 // base tile edge
constexpr int FMAXL = 3;               // levels per launch
constexpr int FE0 = 8 * (FB / 8) + 21; // 85: base extent for three levels; level k extent = (FE0 - 3 * (2^k - 1)) / 2^k

struct FusedArgs {
    const uint8_t* src[2];             // base level of both images (caller's frames, or level `base` of the pyramid)
    int src_pitch[2];
    uint8_t* dst[2];                   // pyramid allocations of both images
    int w[FMAXL + 1], h[FMAXL + 1], pitch[FMAXL + 1];
    size_t off[FMAXL + 1];             // byte offsets of base level and the levels built from it
    int nl;                            // levels built by this launch (0..3)
    int copy_base;                     // also write the base level into dst (it came from outside the pyramid)
};

__device__ __forceinline__ int fused_extent(int k) { return (FE0 - 3 * ((1 << k) - 1)) >> k; }   // 85, 41, 19, 8

__global__ void __launch_bounds__(256)
pyramid_fused_kernel(const FusedArgs A) {
    __shared__ __align__(16) uint8_t s0[FE0 * FE0];
    __shared__ uint8_t s1[41 * 41];
    __shared__ uint8_t s2[19 * 19];
    __shared__ int s_ix[41 * 5];       // reflected source columns / rows of the level being computed, relative to the tile
    __shared__ int s_iy[41 * 5];
    const int img = blockIdx.z, tid = threadIdx.x, L = A.nl;
    const int T = FB >> L;             // tile edge at the top level
    // needed index range of every level, top down: [lo, hi] before reflection
    int lox[FMAXL + 1], hix[FMAXL + 1], loy[FMAXL + 1], hiy[FMAXL + 1];
#pragma unroll
    for (int k = FMAXL; k >= 0; k--) {          // fully unrolled: the arrays stay in registers
        if (k == L) {
            lox[k] = blockIdx.x * T; hix[k] = min(lox[k] + T, A.w[k]) - 1;
            loy[k] = blockIdx.y * T; hiy[k] = min(loy[k] + T, A.h[k]) - 1;
        } else if (k < L) {
            lox[k] = 2 * lox[k + 1 <= FMAXL ? k + 1 : FMAXL] - 2; hix[k] = 2 * hix[k + 1 <= FMAXL ? k + 1 : FMAXL] + 2;
            loy[k] = 2 * loy[k + 1 <= FMAXL ? k + 1 : FMAXL] - 2; hiy[k] = 2 * hiy[k + 1 <= FMAXL ? k + 1 : FMAXL] + 2;
        } else {
            lox[k] = hix[k] = loy[k] = hiy[k] = 0;
        }
    }
    // stored (clipped) range of every level: [c?, c? + n?)
    int cx[FMAXL + 1], cy[FMAXL + 1], nx[FMAXL + 1], ny[FMAXL + 1];
#pragma unroll
    for (int k = 0; k <= FMAXL; k++) {
        cx[k] = max(lox[k], 0); nx[k] = k <= L ? min(hix[k], A.w[k] - 1) - cx[k] + 1 : 0;
        cy[k] = max(loy[k], 0); ny[k] = k <= L ? min(hiy[k], A.h[k] - 1) - cy[k] + 1 : 0;
    }
    // ---- base level: stage, and (first launch on caller-owned frames) copy the owned B x B tile into the pyramid
    {
        const uint8_t* __restrict__ src = A.src[img];
        const int sp = A.src_pitch[img], n = nx[0] * ny[0];
        for (int i = tid; i < n; i += 256) {
            const int y = i / nx[0], x = i - y * nx[0];
            s0[y * FE0 + x] = src[(size_t)(cy[0] + y) * sp + cx[0] + x];
        }
        __syncthreads();
        if (A.copy_base) {
            uint8_t* __restrict__ d = A.dst[img] + A.off[0];
            const int ox0 = blockIdx.x * FB, oy0 = blockIdx.y * FB;
            const int ow = min(FB, A.w[0] - ox0), oh = min(FB, A.h[0] - oy0);
            for (int i = tid; i < ow * oh; i += 256) {
                const int y = i / ow, x = i - y * ow;
                d[(size_t)(oy0 + y) * A.pitch[0] + ox0 + x] = s0[(oy0 + y - cy[0]) * FE0 + (ox0 + x - cx[0])];
            }
        }
    }
    // ---- levels 1..L
#pragma unroll
    for (int k = 1; k <= FMAXL; k++) {
        if (k > L) break;
        const uint8_t* sp = k == 1 ? s0 : (k == 2 ? s1 : s2);
        const int sstr = k == 1 ? FE0 : (k == 2 ? 41 : 19);
        uint8_t* dp = k == 1 ? s1 : (k == 2 ? s2 : nullptr);
        const int dstr = k == 1 ? 41 : 19;
        // reflected source indices (BORDER_REFLECT_101 on the SOURCE level), once per output column / row
        for (int i = tid; i < nx[k] * 5; i += 256) {
            const int x = i / 5, a = i - 5 * x;
            s_ix[i] = reflect101(2 * (cx[k] + x) + a - 2, A.w[k - 1]) - cx[k - 1];
        }
        for (int i = tid; i < ny[k] * 5; i += 256) {
            const int y = i / 5, a = i - 5 * y;
            s_iy[i] = reflect101(2 * (cy[k] + y) + a - 2, A.h[k - 1]) - cy[k - 1];
        }
        __syncthreads();
        uint8_t* __restrict__ d = A.dst[img] + A.off[k];
        const int own = FB >> k;
        const int ox0 = blockIdx.x * own, oy0 = blockIdx.y * own;
        for (int i = tid; i < nx[k] * ny[k]; i += 256) {
            const int y = i / nx[k], x = i - y * nx[k];
            const int* ix = &s_ix[5 * x];
            const int* iy = &s_iy[5 * y];
            int s = 0;
#pragma unroll
            for (int b = 0; b < 5; b++) {
                const uint8_t* r = sp + iy[b] * sstr;
                const int hsum = r[ix[0]] + 4 * r[ix[1]] + 6 * r[ix[2]] + 4 * r[ix[3]] + r[ix[4]];
                s += (b == 0 || b == 4 ? 1 : (b == 2 ? 6 : 4)) * hsum;
            }
            const uint8_t v = (uint8_t)((s + 128) >> 8);
            if (dp) dp[y * dstr + x] = v;
            const int gx = cx[k] + x, gy = cy[k] + y;
            if (gx >= ox0 && gx < ox0 + own && gy >= oy0 && gy < oy0 + own) d[(size_t)gy * A.pitch[k] + gx] = v;
        }
        __syncthreads();
    }
}

int layout_pyramid(fm3d_ctx* ctx, int w, int h, int levels) {
    fm3d_pyramid_desc& P = ctx->pyr;
    size_t off = 0;
    for (int l = 0; l <= levels; l++) {
        P.lv[l].w = w; P.lv[l].h = h;
        P.lv[l].pitch = (w + 15) & ~15;
        P.lv[l].off = off;
        off += ((size_t)P.lv[l].pitch * h + 255) & ~(size_t)255;
        w = (w + 1) / 2; h = (h + 1) / 2;
    }
    P.levels = levels;
    const size_t per_image = off;
    if (ctx->pyr_bytes < 2 * per_image) {
        if (ctx->pyr_mem) {
            FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            FM3D_CUDA(ctx, cudaFree(ctx->pyr_mem));
            ctx->pyr_mem = nullptr; ctx->pyr_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->pyr_mem, 2 * per_image);
        if (e != cudaSuccess) return fm3d_fail(ctx, FM3D_ERR_NOMEM, "cudaMalloc(pyramids) failed: %s", cudaGetErrorString(e));
        ctx->pyr_bytes = 2 * per_image;
    }
    P.base[0] = ctx->pyr_mem;
    P.base[1] = ctx->pyr_mem + per_image;
    return FM3D_OK;
}

// Builds levels 1..L (and level 0 from img1 / img2 when those are device pointers outside the pyramid) with the fused
// kernel: one launch per three levels.
int build_levels_fused(fm3d_ctx* ctx, const uint8_t* img1, int stride1, const uint8_t* img2, int stride2) {
    fm3d_pyramid_desc& P = ctx->pyr;
    uint8_t* d0 = const_cast<uint8_t*>(P.base[0]);
    uint8_t* d1 = const_cast<uint8_t*>(P.base[1]);
    int base = 0;
    bool first = true;
    do {
        FusedArgs A{};
        const int nl = P.levels - base < FMAXL ? P.levels - base : FMAXL;
        const bool ext = first && img1 != nullptr;
        A.src[0] = ext ? img1 : d0 + P.lv[base].off; A.src_pitch[0] = ext ? stride1 : P.lv[base].pitch;
        A.src[1] = ext ? img2 : d1 + P.lv[base].off; A.src_pitch[1] = ext ? stride2 : P.lv[base].pitch;
        A.dst[0] = d0; A.dst[1] = d1;
        for (int k = 0; k <= nl; k++) {
            const fm3d_level& lv = P.lv[base + k];
            A.w[k] = lv.w; A.h[k] = lv.h; A.pitch[k] = lv.pitch; A.off[k] = lv.off;
        }
        A.nl = nl;
        A.copy_base = ext ? 1 : 0;
        if (nl > 0 || A.copy_base) {
            const int T = FB >> nl;
            dim3 grid((A.w[nl] + T - 1) / T, (A.h[nl] + T - 1) / T, 2);
            pyramid_fused_kernel<<<grid, 256, 0, ctx->stream>>>(A);
            FM3D_LAUNCH_CHECK(ctx);
        }
        base += nl;
        first = false;
    } while (base < P.levels);
    ctx->has_images = true;
    return FM3D_OK;
}

int build_levels(fm3d_ctx* ctx) {
    fm3d_pyramid_desc& P = ctx->pyr;
    // rows are pitched to 16 bytes; the pad bytes of level 0 are whatever cudaMalloc left there and those of the other
    // levels whatever the 32-bit tile stores put there: every reader (this kernel, the TMA tensor maps, the samplers) is
    // bounded by the level's width, so they are never read
    for (int l = 1; l <= P.levels; l++) {
        const fm3d_level& s = P.lv[l - 1];
        const fm3d_level& d = P.lv[l];
        dim3 grid((d.w + TW - 1) / TW, (d.h + TH - 1) / TH, 2);
        pyrdown_kernel<<<grid, 256, 0, ctx->stream>>>(P.base[0] + s.off, P.base[1] + s.off, s.w, s.h, s.pitch,
                                                      const_cast<uint8_t*>(P.base[0]) + d.off,
                                                      const_cast<uint8_t*>(P.base[1]) + d.off, d.w, d.h, d.pitch);
        FM3D_LAUNCH_CHECK(ctx);
    }
    ctx->has_images = true;
    return FM3D_OK;
}

int set_images_common(fm3d_ctx* ctx, const uint8_t* img1, int stride1, const uint8_t* img2, int stride2, int w, int h,
                      int pyramids, cudaMemcpyKind kind) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, img1 && img2 && w > 0 && h > 0 && stride1 >= w && stride2 >= w);
    FM3D_CHECK_ARG(ctx, pyramids >= 0 && pyramids < FM3D_MAX_LEVELS);
    if (int rc = fm3d_bind(ctx)) return rc;
    ctx->has_images = false;
    if (int rc = layout_pyramid(ctx, w, h, pyramids)) return rc;
    const fm3d_level& l0 = ctx->pyr.lv[0];
    if (ctx->opt_pyramid_fused) {
        // device frames: the fused kernel reads them where they are and writes level 0 itself; host frames: one upload per
        // frame straight into level 0, then the same kernel
        if (kind == cudaMemcpyDeviceToDevice) return build_levels_fused(ctx, img1, stride1, img2, stride2);
        for (int k = 0; k < 2; k++) {
            uint8_t* d = const_cast<uint8_t*>(ctx->pyr.base[k]) + l0.off;
            ctx->n_copy++;
            FM3D_CUDA(ctx, cudaMemcpy2DAsync(d, l0.pitch, k ? img2 : img1, k ? stride2 : stride1, w, h, kind, ctx->stream));
        }
        return build_levels_fused(ctx, nullptr, 0, nullptr, 0);
    }
    for (int k = 0; k < 2; k++) {
        uint8_t* d = const_cast<uint8_t*>(ctx->pyr.base[k]) + l0.off;
        ctx->n_copy++;
        FM3D_CUDA(ctx, cudaMemcpy2DAsync(d, l0.pitch, k ? img2 : img1, k ? stride2 : stride1, w, h, kind, ctx->stream));
    }
    return build_levels(ctx);
}

}  // namespace

extern "C" {

int fm3d_set_images(fm3d_ctx* ctx, const uint8_t* img1, const uint8_t* img2, int w, int h, int stride,
                    int pyramids) {
    int rc = set_images_common(ctx, img1, stride, img2, stride, w, h, pyramids, cudaMemcpyHostToDevice);
    if (rc) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_set_images2(fm3d_ctx* ctx, const uint8_t* img1, int stride1, const uint8_t* img2, int stride2, int w, int h,
                     int pyramids) {
    int rc = set_images_common(ctx, img1, stride1, img2, stride2, w, h, pyramids, cudaMemcpyHostToDevice);
    if (rc) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_set_images_dev(fm3d_ctx* ctx, const uint8_t* img1, const uint8_t* img2, int w, int h,
                        int stride, int pyramids) {
    return set_images_common(ctx, img1, stride, img2, stride, w, h, pyramids, cudaMemcpyDeviceToDevice);
}

int fm3d_get_pyramid_level(fm3d_ctx* ctx, int image, int level, uint8_t* out, int* w, int* h) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    FM3D_CHECK_ARG(ctx, (image == 1 || image == 2) && level >= 0 && level <= ctx->pyr.levels);
    if (int rc = fm3d_bind(ctx)) return rc;
    const fm3d_level& L = ctx->pyr.lv[level];
    if (w) *w = L.w;
    if (h) *h = L.h;
    if (out) {
        ctx->n_copy++;
        FM3D_CUDA(ctx, cudaMemcpy2DAsync(out, L.w, ctx->pyr.base[image - 1] + L.off, L.pitch, L.w, L.h,
                                         cudaMemcpyDeviceToHost, ctx->stream));
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FM3D_OK;
}

}  // extern "C"
