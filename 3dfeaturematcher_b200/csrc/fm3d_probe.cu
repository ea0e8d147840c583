// fm3d_probe.cu -- the per-evaluation helpers of SingleCameraTriangulator as stand-alone entry points.
//
// The reference's NormalOptimizer::evaluateNormal (Triangulator/normaloptimizer.cpp:65-149) is built
// from four PUBLIC methods of SingleCameraTriangulator:
//   extractPixelsContour(Vec3d)          singlecameratriangulator.cpp:341-397   disc lattice
//   get3dPointsFromImage1Pixels          :530-565 (+ projectPointToPlane :421-470, isInBoundingBox :646-655)
//   updateImage1PixelsIntensity          :576-589 (+ isPixelGood :657-665, getBilinearInterpPix32f tools.cpp:129-142)
//   projectPointsToImage2                :591-632
// libfm3d fuses them into the normal-search kernels; a caller of the class interface can still
// call them one by one, so each has an element-wise kernel here (one thread per pixel, fp64
// geometry, the reference's fp32 sampler).  They double as per-pixel parity probes of the fused
// kernels.  Not a hot path: host-pointer entry points only.
#include "fm3d_internal.cuh"

namespace {

enum { PROBE_BBOX = 1, PROBE_NAN = 2, PROBE_PIXEL = 4 };

// extractPixelsContour(Vec2d): for i = -r..r (x offset, OUTER loop), j = -r..r, keep (cx+i, cy+j) if
// i^2+j^2 <= r^2 and the pixel is inside the image (the reference hard-codes 1024x768: D1).
__global__ void disc_count_kernel(double cx, double cy, int r, int W, int H, int* __restrict__ col_count) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > 2 * r) return;
    const int i = c - r;
    const double px = cx + (double)i;
    int cnt = 0;
    if (!(px < 0 || px >= (double)W) && cx == cx && cy == cy) {
        for (int j = -r; j <= r; j++) {
            if (i * i + j * j > r * r) continue;
            const double py = cy + (double)j;
            if (!(py < 0 || py >= (double)H)) cnt++;
        }
    }
    col_count[c] = cnt;
}

__global__ void disc_fill_kernel(double cx, double cy, int r, int W, int H, const int* __restrict__ col_start,
                                 double* __restrict__ xy, int cap) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > 2 * r) return;
    const int i = c - r;
    const double px = cx + (double)i;
    if (px < 0 || px >= (double)W || cx != cx || cy != cy) return;
    int o = col_start[c];
    for (int j = -r; j <= r; j++) {
        if (i * i + j * j > r * r) continue;
        const double py = cy + (double)j;
        if (py < 0 || py >= (double)H) continue;
        if (o < cap) { xy[2 * o] = px; xy[2 * o + 1] = py; }
        o++;
    }
}

__global__ void scan_columns_kernel(int* __restrict__ counts, int n, int* __restrict__ total) {
    // n <= 511: one thread is enough
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int acc = 0;
    for (int k = 0; k < n; k++) { const int c = counts[k]; counts[k] = acc; acc += c; }
    *total = acc;
}

// get3dPointsFromImage1Pixels: u = undistortPoints(pixel), v = (u, 1), X = ((n.P)/(n.v)) v
__global__ void plane_points_kernel(fm3d_cam cam, double Px, double Py, double Pz, double nx, double ny, double nz,
                                    const double* __restrict__ xy, int m, double* __restrict__ xyz, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double vx, vy;
    fm3d_undistort(cam, xy[2 * i], xy[2 * i + 1], vx, vy);
    const double k = (nx * Px + ny * Py + nz * Pz) / (nx * vx + ny * vy + nz);   // projectPointToPlane (:421-470)
    const double X = k * vx, Y = k * vy, Z = k;
    xyz[3 * i] = X; xyz[3 * i + 1] = Y; xyz[3 * i + 2] = Z;
    int f = 0;
    if (X != X || Y != Y || Z != Z) f |= PROBE_NAN;
    const double cmax = (double)(int)(2 * cam.zmax);                             // isInBoundingBox (:646-655)
    if (!((X > -cmax && X < cmax) && (Y > -cmax && Y < cmax) && (Z > 0 && Z < cmax))) f |= PROBE_BBOX;
    if (f) atomicOr(flags, f);
}

// updateImage1PixelsIntensity / the sampling half of projectPointsToImage2:
// gate = isPixelGood(p, scale) on the level's cols/rows, value = bilinear(img, (float)(scale x), (float)(scale y))
__global__ void sample_kernel(const uint8_t* __restrict__ img, int w, int h, int pitch, double scale, int gate,
                              const double* __restrict__ xy, int m, float* __restrict__ out, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double x = xy[2 * i], y = xy[2 * i + 1];
    if (gate && !fm3d_pixel_good(x, y, 1.0 / scale, w, h)) atomicOr(flags, PROBE_PIXEL);
    out[i] = fm3d_bilinear_global(img, w, h, pitch, (float)(scale * x), (float)(scale * y));
}

// projectPointsToImage2: cv::projectPoints(X, r2, t2 = g12, K, dist)
__global__ void project2_kernel(fm3d_cam cam, const double* __restrict__ xyz, int m, double* __restrict__ xy2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const double X = xyz[3 * i], Y = xyz[3 * i + 1], Z = xyz[3 * i + 2];
    const double X2 = cam.R[0] * X + cam.R[1] * Y + cam.R[2] * Z + cam.t[0];
    const double Y2 = cam.R[3] * X + cam.R[4] * Y + cam.R[5] * Z + cam.t[1];
    const double Z2 = cam.R[6] * X + cam.R[7] * Y + cam.R[8] * Z + cam.t[2];
    double u, v;
    fm3d_project(cam, X2, Y2, Z2, u, v);
    xy2[2 * i] = u; xy2[2 * i + 1] = v;
}

int info_from_flags(int f) {
    if (f & PROBE_NAN) return -6;      // the reference exit(-6)s (singlecameratriangulator.cpp:465-469)
    if (f & (PROBE_BBOX | PROBE_PIXEL)) return -1;
    return 0;
}

}  // namespace

extern "C" {

int fm3d_disc_pixels(fm3d_ctx* ctx, const double P[3], int pixels_ray, double* xy, int cap, int* m) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, P && m && cap >= 0 && (cap == 0 || xy) && pixels_ray >= 0 && pixels_ray <= 255);
    if (!ctx->has_cam) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera not set");
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set (the disc is clipped to the image size)");
    if (int rc = fm3d_bind(ctx)) return rc;
    const fm3d_cam& c = ctx->cam;
    // extractPixelsContour(Vec3d): centre = cv::projectPoints(P, 0, 0, K, dist) (:376-397), host arithmetic
    const double z = P[2] != 0.0 ? 1.0 / P[2] : 1.0;
    const double x = P[0] * z, y = P[1] * z;
    const double r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
    const double a1 = 2 * x * y, a2 = r2 + 2 * x * x, a3 = r2 + 2 * y * y;
    const double cd = 1 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
    const double cu = (x * cd + c.p1 * a1 + c.p2 * a2) * c.fx + c.cx, cv = (y * cd + c.p1 * a3 + c.p2 * a1) * c.fy + c.cy;
    const int ncol = 2 * pixels_ray + 1, W = ctx->pyr.lv[0].w, H = ctx->pyr.lv[0].h;
    char* d = nullptr;
    const size_t o_xy = 4096, bytes = o_xy + sizeof(double) * 2 * (size_t)cap;
    if (int rc = fm3d_scratch(ctx, 0, bytes, (void**)&d)) return rc;
    int* counts = (int*)d;
    int* total = counts + 600;
    disc_count_kernel<<<(ncol + 127) / 128, 128, 0, ctx->stream>>>(cu, cv, pixels_ray, W, H, counts);
    FM3D_LAUNCH_CHECK(ctx);
    scan_columns_kernel<<<1, 32, 0, ctx->stream>>>(counts, ncol, total);
    FM3D_LAUNCH_CHECK(ctx);
    disc_fill_kernel<<<(ncol + 127) / 128, 128, 0, ctx->stream>>>(cu, cv, pixels_ray, W, H, counts, (double*)(d + o_xy), cap);
    FM3D_LAUNCH_CHECK(ctx);
    int tot = 0;
    if (int rc = fm3d_d2h(ctx, &tot, total, sizeof(int))) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *m = tot;
    const int ncopy = tot < cap ? tot : cap;
    if (ncopy > 0) {
        if (int rc = fm3d_d2h(ctx, xy, d + o_xy, sizeof(double) * 2 * (size_t)ncopy)) return rc;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return FM3D_OK;
}

int fm3d_plane_points(fm3d_ctx* ctx, const double P[3], const double normal[3], const double* xy, int m,
                      double* xyz, int* info) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, P && normal && info && m >= 0 && (m == 0 || (xy && xyz)));
    if (!ctx->has_cam) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera not set");
    *info = 0;
    if (m == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b2 = sizeof(double) * 2 * (size_t)m, b3 = sizeof(double) * 3 * (size_t)m;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 256 + al(b2) + al(b3), (void**)&d)) return rc;
    int* flags = (int*)d;
    ctx->n_copy++;
    FM3D_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int), ctx->stream));
    if (int rc = fm3d_h2d(ctx, d + 256, xy, b2)) return rc;
    plane_points_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->cam, P[0], P[1], P[2], normal[0], normal[1], normal[2],
                                                                  (const double*)(d + 256), m, (double*)(d + 256 + al(b2)), flags);
    FM3D_LAUNCH_CHECK(ctx);
    int f = 0;
    if (int rc = fm3d_d2h(ctx, xyz, d + 256 + al(b2), b3)) return rc;
    if (int rc = fm3d_d2h(ctx, &f, flags, sizeof(int))) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = info_from_flags(f);
    return FM3D_OK;
}

int fm3d_sample_pixels(fm3d_ctx* ctx, int image, int level, double scale, int gate, const double* xy, int m,
                       float* intensity, int* info) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, info && (image == 1 || image == 2) && scale > 0 && m >= 0 && (m == 0 || (xy && intensity)));
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    FM3D_CHECK_ARG(ctx, level >= 0 && level <= ctx->pyr.levels);
    *info = 0;
    if (m == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b2 = sizeof(double) * 2 * (size_t)m, b1 = sizeof(float) * (size_t)m;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 256 + al(b2) + al(b1), (void**)&d)) return rc;
    int* flags = (int*)d;
    ctx->n_copy++;
    FM3D_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int), ctx->stream));
    if (int rc = fm3d_h2d(ctx, d + 256, xy, b2)) return rc;
    const fm3d_level& lv = ctx->pyr.lv[level];
    sample_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->pyr.base[image - 1] + lv.off, lv.w, lv.h, lv.pitch, scale, gate,
                                                            (const double*)(d + 256), m, (float*)(d + 256 + al(b2)), flags);
    FM3D_LAUNCH_CHECK(ctx);
    int f = 0;
    if (int rc = fm3d_d2h(ctx, intensity, d + 256 + al(b2), b1)) return rc;
    if (int rc = fm3d_d2h(ctx, &f, flags, sizeof(int))) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = info_from_flags(f);
    return FM3D_OK;
}

int fm3d_project_to_image2(fm3d_ctx* ctx, const double* xyz, int m, int level, double scale, double* xy2,
                           float* intensity, int* info) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, info && scale > 0 && m >= 0 && (m == 0 || (xyz && xy2)));
    if (!ctx->has_cam || !ctx->has_g12) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera / g12 not set");
    if (intensity && !ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    if (intensity) FM3D_CHECK_ARG(ctx, level >= 0 && level <= ctx->pyr.levels);
    *info = 0;
    if (m == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b3 = sizeof(double) * 3 * (size_t)m, b2 = sizeof(double) * 2 * (size_t)m, b1 = sizeof(float) * (size_t)m;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 256 + al(b3) + al(b2) + al(b1), (void**)&d)) return rc;
    int* flags = (int*)d;
    ctx->n_copy++;
    FM3D_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int), ctx->stream));
    if (int rc = fm3d_h2d(ctx, d + 256, xyz, b3)) return rc;
    double* d_xy2 = (double*)(d + 256 + al(b3));
    project2_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->cam, (const double*)(d + 256), m, d_xy2);
    FM3D_LAUNCH_CHECK(ctx);
    if (intensity) {
        const fm3d_level& lv = ctx->pyr.lv[level];
        sample_kernel<<<(m + 255) / 256, 256, 0, ctx->stream>>>(ctx->pyr.base[1] + lv.off, lv.w, lv.h, lv.pitch, scale, 1, d_xy2, m,
                                                                (float*)(d + 256 + al(b3) + al(b2)), flags);
        FM3D_LAUNCH_CHECK(ctx);
        if (int rc = fm3d_d2h(ctx, intensity, d + 256 + al(b3) + al(b2), b1)) return rc;
    }
    int f = 0;
    if (int rc = fm3d_d2h(ctx, xy2, d_xy2, b2)) return rc;
    if (int rc = fm3d_d2h(ctx, &f, flags, sizeof(int))) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = info_from_flags(f);
    return FM3D_OK;
}

}  // extern "C"
