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
