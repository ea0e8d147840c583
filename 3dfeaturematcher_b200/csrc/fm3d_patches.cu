// fm3d_patches.cu -- K8: rectified patch extraction.
//
// Replaces NeighborhoodsGenerator::getReferenceSquaredNeighborhood
// (Triangulator/neighborhoodsgenerator.cpp:134-158) + SingleCameraTriangulator::
// projectReferencePointsToImageWithFrames (Triangulator/singlecameratriangulator.cpp:769-849),
// the explicit-group variant projectPointsToImage (:667-767) and
// computeSquareNeighborhoodsByNormals (neighborhoodsgenerator.cpp:76-132).
//
// The S x S metric grid is generated analytically (the reference materialises 393 KB of
// reference points and 262 KB of image points per feature).  One CTA handles one 32x32 tile
// of one patch: threads run with the grid's i index fastest so that the u8 patch rows
// (patch(row=j, col=i), the reference writes patch.at<uchar>(col,row)) are stored as
// coalesced bytes; the optional image points, whose order is idx = i*S + j, are transposed
// through shared memory and stored as coalesced double2.  Geometry in fp64, sampling in fp32
// exactly as getBilinearInterpPix32f.  HBM-bound on the 16 B/pixel image-point stream when it
// is requested, otherwise gather/ALU-bound.
//
// The reference round-trips the frame's rotation through cv::Rodrigues
// (decomposeTransformation, tools.cpp:101-114, then cv::projectPoints); for the orthonormal
// frames computeFeaturesFrames produces that is the identity to 1e-16, so R is used directly.
#include "fm3d_internal.cuh"

namespace {

constexpr int PT = 32;  // tile edge

struct PatchArgs {
    fm3d_cam cam;
    const uint8_t* img;
    int w, h, pitch;
    int S;
    double eps_m, inc;
    int use_g12;            // project_groups with image 2
};

__device__ __forceinline__ uint8_t sample_patch(const PatchArgs& A, double u, double v) {
    // isPixelGood(p, 1.0) else 0 (:838-847); static_cast<uchar> truncates
    if (!fm3d_pixel_good(u, v, 1.0, A.w, A.h) || u != u || v != v) return 0;
    const float val = fm3d_bilinear_global(A.img, A.w, A.h, A.pitch, (float)u, (float)v);
    return (uint8_t)val;
}

// frames: n x 16.  grid = (tiles_i, tiles_j, n)
__global__ void __launch_bounds__(256)
patches_kernel(PatchArgs A, const double* __restrict__ frames, uint8_t* __restrict__ patches,
               double* __restrict__ image_points) {
    __shared__ double2 s_ip[PT][PT + 1];
    const int f = blockIdx.z;
    const int i0 = blockIdx.x * PT, j0 = blockIdx.y * PT;
    const double* F = frames + 16 * (size_t)f;
    const double r00 = F[0], r01 = F[1], t0 = F[3];
    const double r10 = F[4], r11 = F[5], t1 = F[7];
    const double r20 = F[8], r21 = F[9], t2 = F[11];
    const int S = A.S;
    const int ti = threadIdx.x & 31, tj0 = threadIdx.x >> 5;  // 8 rows of 32
    for (int tj = tj0; tj < PT; tj += 8) {
        const int i = i0 + ti, j = j0 + tj;
        double u = 0, v = 0;
        if (i < S && j < S) {
            const double rx = -A.eps_m + A.inc * (double)i, ry = -A.eps_m + A.inc * (double)j;  // (:150-152)
            const double X = r00 * rx + r01 * ry + t0;
            const double Y = r10 * rx + r11 * ry + t1;
            const double Z = r20 * rx + r21 * ry + t2;
            fm3d_project(A.cam, X, Y, Z, u, v);
            patches[(size_t)f * S * S + (size_t)j * S + i] = sample_patch(A, u, v);
        }
        if (image_points) s_ip[tj][ti] = make_double2(u, v);
    }
    if (image_points) {
        __syncthreads();
        // transposed store: consecutive threads -> consecutive j for fixed i
        const int sj = threadIdx.x & 31, si0 = threadIdx.x >> 5;
        for (int si = si0; si < PT; si += 8) {
            const int i = i0 + si, j = j0 + sj;
            if (i < S && j < S)
                reinterpret_cast<double2*>(image_points)[(size_t)f * S * S + (size_t)i * S + j] = s_ip[sj][si];
        }
    }
}

// groups: n x S*S x 3 explicit points (camera-1 coordinates)
__global__ void __launch_bounds__(256)
groups_kernel(PatchArgs A, const double* __restrict__ groups, uint8_t* __restrict__ patches,
              double* __restrict__ image_points) {
    const int S = A.S;
    const size_t per = (size_t)S * S;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (gid >= per) return;
    const double* p = groups + ((size_t)f * per + gid) * 3;
    double X = p[0], Y = p[1], Z = p[2];
    if (A.use_g12) {
        const fm3d_cam& c = A.cam;
        const double x2 = c.R[0] * X + c.R[1] * Y + c.R[2] * Z + c.t[0];
        const double y2 = c.R[3] * X + c.R[4] * Y + c.R[5] * Z + c.t[1];
        const double z2 = c.R[6] * X + c.R[7] * Y + c.R[8] * Z + c.t[2];
        X = x2; Y = y2; Z = z2;
    }
    double u, v;
    fm3d_project(A.cam, X, Y, Z, u, v);
    const int col = (int)(gid % S), row = (int)(gid / S);   // (:743-747)
    patches[(size_t)f * per + (size_t)col * S + row] = sample_patch(A, u, v);
    if (image_points) reinterpret_cast<double2*>(image_points)[(size_t)f * per + gid] = make_double2(u, v);
}

__global__ void __launch_bounds__(256)
neighborhoods_kernel(const double* __restrict__ frames, int S, double eps_m, double inc,
                     double* __restrict__ out) {
    const size_t per = (size_t)S * S;
    const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (gid >= per) return;
    const double* F = frames + 16 * (size_t)f;
    const int i = (int)(gid / S), j = (int)(gid % S);
    const double rx = -eps_m + inc * (double)i, ry = -eps_m + inc * (double)j;
    double X = F[0] * rx + F[1] * ry + F[3];
    double Y = F[4] * rx + F[5] * ry + F[7];
    double Z = F[8] * rx + F[9] * ry + F[11];
    const double Wh = F[12] * rx + F[13] * ry + F[15];
    if (Wh != 1.0) { X /= Wh; Y /= Wh; Z /= Wh; }  // neighborhoodsgenerator.cpp:119-122
    double* o = out + ((size_t)f * per + gid) * 3;
    o[0] = X; o[1] = Y; o[2] = Z;
}

int patch_args(fm3d_ctx* ctx, PatchArgs& A, int image, double eps_m, double cm_per_pixel, int S) {
    if (!ctx->has_cam) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera not set");
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    const fm3d_level& l0 = ctx->pyr.lv[0];
    A.cam = ctx->cam;
    A.img = ctx->pyr.base[image - 1] + l0.off;
    A.w = l0.w; A.h = l0.h; A.pitch = l0.pitch;
    A.S = S; A.eps_m = eps_m; A.inc = cm_per_pixel * 0.01;
    A.use_g12 = image == 2;
    return FM3D_OK;
}

// NeighborhoodsGenerator::computeCircularNeighborhood(s)ByNormal(s) (neighborhoodsgenerator.cpp:160-277):
// samples on concentric circles of the feature plane by Rodrigues' rotation of a spanner vector,
//   sample k = (ray i in 1..n_rays, angle j):  r = i eps / n_rays, theta = j 2 pi / n_angles,
//   s = (0, 1, -ny/nz) / |.| * eps,   p_k = P + r (s + sin(theta) n x s + 2 sin^2(theta/2) n x (n x s))
// (the radius ends up scaled by eps twice: as in the reference).  An all-zero normal is replaced
// by P/|P| and written back.
__global__ void circular_kernel(const double* __restrict__ pts, double* __restrict__ normals, int n, double eps,
                                int n_angles, int n_rays, double* __restrict__ out) {
    const int S = n_angles * n_rays;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * S) return;
    const int f = (int)(gid / S), k = (int)(gid - (long long)f * S);
    const double Px = pts[3 * f], Py = pts[3 * f + 1], Pz = pts[3 * f + 2];
    double nx = normals[3 * f], ny = normals[3 * f + 1], nz = normals[3 * f + 2];
    if (nx == 0 && ny == 0 && nz == 0) {
        const double nr = sqrt(Px * Px + Py * Py + Pz * Pz);
        nx = Px / nr; ny = Py / nr; nz = Pz / nr;
    }
    const int i = k / n_angles + 1, j = k - (k / n_angles) * n_angles;
    const double r = (double)i * (eps / (double)n_rays);
    const double theta = (double)j * (2 * 3.14159265358979323846 / (double)n_angles);
    const double st = sin(theta), sh = sin(theta / 2), st2 = 2 * sh * sh;
    double sx = 0.0, sy = 1.0, sz = -ny / nz;                   // <spanner, normal> = 0
    const double sn = sqrt(sx * sx + sy * sy + sz * sz);
    sx = sx / sn * eps; sy = sy / sn * eps; sz = sz / sn * eps;
    const double ax = ny * sz - nz * sy, ay = nz * sx - nx * sz, az = nx * sy - ny * sx;      // W s  = n x s
    const double bx = ny * az - nz * ay, by = nz * ax - nx * az, bz = nx * ay - ny * ax;      // W W s
    double* o = out + 3 * (size_t)gid;
    o[0] = Px + r * (sx + ax * st + st2 * bx);
    o[1] = Py + r * (sy + ay * st + st2 * by);
    o[2] = Pz + r * (sz + az * st + st2 * bz);
}

__global__ void fill_default_normals_kernel(const double* __restrict__ pts, double* __restrict__ normals, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    if (normals[3 * f] == 0 && normals[3 * f + 1] == 0 && normals[3 * f + 2] == 0) {
        const double Px = pts[3 * f], Py = pts[3 * f + 1], Pz = pts[3 * f + 2];
        const double nr = sqrt(Px * Px + Py * Py + Pz * Pz);
        normals[3 * f] = Px / nr; normals[3 * f + 1] = Py / nr; normals[3 * f + 2] = Pz / nr;
    }
}

}  // namespace

extern "C" {

int fm3d_patch_size(double epsilon_m, double cm_per_pixel) {
    if (!(cm_per_pixel > 0) || !(epsilon_m >= 0)) return 0;
    return 2 * ((int)floor(epsilon_m / (0.01 * cm_per_pixel)));
}

int fm3d_extract_patches_dev(fm3d_ctx* ctx, const double* frames, int n, double epsilon_m,
                             double cm_per_pixel, uint8_t* patches, double* image_points) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    const int S = fm3d_patch_size(epsilon_m, cm_per_pixel);
    FM3D_CHECK_ARG(ctx, n >= 0 && S > 0 && (n == 0 || (frames && patches)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    PatchArgs A;
    if (int rc = patch_args(ctx, A, 1, epsilon_m, cm_per_pixel, S)) return rc;
    const int tiles = (S + PT - 1) / PT;
    for (int f0 = 0; f0 < n; f0 += 65535) {
        const int nf = n - f0 < 65535 ? n - f0 : 65535;
        dim3 grid(tiles, tiles, nf);
        patches_kernel<<<grid, 256, 0, ctx->stream>>>(A, frames + 16 * (size_t)f0, patches + (size_t)f0 * S * S,
                                                      image_points ? image_points + (size_t)f0 * S * S * 2 : nullptr);
        FM3D_LAUNCH_CHECK(ctx);
    }
    return FM3D_OK;
}

int fm3d_extract_patches(fm3d_ctx* ctx, const double* frames, int n, double epsilon_m,
                         double cm_per_pixel, uint8_t* patches, double* image_points) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    const int S = fm3d_patch_size(epsilon_m, cm_per_pixel);
    FM3D_CHECK_ARG(ctx, n >= 0 && S > 0 && (n == 0 || (frames && patches)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bf = sizeof(double) * 16 * (size_t)n, bp = (size_t)n * S * S;
    const size_t bip = image_points ? sizeof(double) * 2 * bp : 0;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bf) + al(bp) + al(bip), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, frames, bf)) return rc;
    double* d_ip = image_points ? (double*)(d + al(bf) + al(bp)) : nullptr;
    if (int rc = fm3d_extract_patches_dev(ctx, (const double*)d, n, epsilon_m, cm_per_pixel,
                                          (uint8_t*)(d + al(bf)), d_ip)) return rc;
    if (int rc = fm3d_d2h(ctx, patches, d + al(bf), bp)) return rc;
    if (image_points) if (int rc = fm3d_d2h(ctx, image_points, d_ip, bip)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_project_groups(fm3d_ctx* ctx, int image, const double* groups, int n, int S,
                        uint8_t* patches, double* image_points) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, (image == 1 || image == 2) && n >= 0 && n <= 65535 && S > 0 && (n == 0 || (groups && patches)));
    if (n == 0) return FM3D_OK;
    if (image == 2 && !ctx->has_g12) return fm3d_fail(ctx, FM3D_ERR_STATE, "g12 not set");
    if (int rc = fm3d_bind(ctx)) return rc;
    PatchArgs A;
    if (int rc = patch_args(ctx, A, image, 0, 0, S)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t per = (size_t)S * S, bg = sizeof(double) * 3 * per * n, bp = per * n;
    const size_t bip = image_points ? sizeof(double) * 2 * per * n : 0;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bg) + al(bp) + al(bip), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, groups, bg)) return rc;
    double* d_ip = image_points ? (double*)(d + al(bg) + al(bp)) : nullptr;
    dim3 grid((unsigned)((per + 255) / 256), n);
    groups_kernel<<<grid, 256, 0, ctx->stream>>>(A, (const double*)d, (uint8_t*)(d + al(bg)), d_ip);
    FM3D_LAUNCH_CHECK(ctx);
    if (int rc = fm3d_d2h(ctx, patches, d + al(bg), bp)) return rc;
    if (image_points) if (int rc = fm3d_d2h(ctx, image_points, d_ip, bip)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_square_neighborhoods(fm3d_ctx* ctx, const double* frames, int n, double epsilon_m,
                              double cm_per_pixel, double* out) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    const int S = fm3d_patch_size(epsilon_m, cm_per_pixel);
    FM3D_CHECK_ARG(ctx, n >= 0 && n <= 65535 && S > 0 && (n == 0 || (frames && out)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t per = (size_t)S * S, bf = sizeof(double) * 16 * (size_t)n, bo = sizeof(double) * 3 * per * n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bf) + al(bo), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, frames, bf)) return rc;
    dim3 grid((unsigned)((per + 255) / 256), n);
    neighborhoods_kernel<<<grid, 256, 0, ctx->stream>>>((const double*)d, S, epsilon_m, cm_per_pixel * 0.01,
                                                        (double*)(d + al(bf)));
    FM3D_LAUNCH_CHECK(ctx);
    if (int rc = fm3d_d2h(ctx, out, d + al(bf), bo)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_circular_neighborhoods(fm3d_ctx* ctx, const double* points, double* normals, int n, double epsilon_m,
                                int n_angles, int n_rays, double* out) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && n_angles >= 1 && n_rays >= 1 && (n == 0 || (points && normals && out)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t S = (size_t)n_angles * n_rays;
    const size_t b3 = sizeof(double) * 3 * (size_t)n, bo = sizeof(double) * 3 * S * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 2 * al(b3) + al(bo), (void**)&d)) return rc;
    double* d_p = (double*)d; double* d_n = (double*)(d + al(b3)); double* d_o = (double*)(d + 2 * al(b3));
    if (int rc = fm3d_h2d(ctx, d_p, points, b3)) return rc;
    if (int rc = fm3d_h2d(ctx, d_n, normals, b3)) return rc;
    const long long total = (long long)n * (long long)S;
    circular_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(d_p, d_n, n, epsilon_m, n_angles, n_rays, d_o);
    FM3D_LAUNCH_CHECK(ctx);
    fill_default_normals_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_p, d_n, n);
    FM3D_LAUNCH_CHECK(ctx);
    if (int rc = fm3d_d2h(ctx, out, d_o, bo)) return rc;
    if (int rc = fm3d_d2h(ctx, normals, d_n, b3)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

}  // extern "C"
