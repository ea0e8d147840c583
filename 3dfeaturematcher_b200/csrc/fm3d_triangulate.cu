// fm3d_triangulate.cu -- K3: gather matched keypoints, undistort, 4x4 DLT, depth gate,
// order-preserving compaction.  Replaces SingleCameraTriangulator::setKeypoints + ::triangulate
// (Triangulator/singlecameratriangulator.cpp:145-230).
//
// One thread per match, fp64 throughout (the reference works on CV_64F).  The null vector of
// the 4x4 DLT system is found with a one-sided Jacobi (Hestenes) SVD kept in registers, the
// method cv::SVD uses for small matrices.  Launch-latency bound: 57 B per match.
#include "fm3d_internal.cuh"

namespace {

__device__ __forceinline__ void null_vector4(double (&U)[4][4], double (&x)[4]) {
    double V[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        bool changed = false;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                double a = 0, b = 0, g = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    a += U[k][p] * U[k][p];
                    b += U[k][q] * U[k][q];
                    g += U[k][p] * U[k][q];
                }
                if (fabs(g) <= 1e-300 || fabs(g) <= 2.220446049250313e-16 * sqrt(a * b)) continue;
                changed = true;
                const double zeta = (b - a) / (2 * g);
                const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
                const double cs = 1 / sqrt(1 + tt * tt), sn = cs * tt;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double up = U[k][p], uq = U[k][q];
                    U[k][p] = cs * up - sn * uq; U[k][q] = sn * up + cs * uq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = cs * vp - sn * vq; V[k][q] = sn * vp + cs * vq;
                }
            }
        if (!changed) break;
    }
    double sm = INFINITY;
    x[0] = x[1] = x[2] = x[3] = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double s = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) s += U[k][j] * U[k][j];
        if (s < sm) {
            sm = s;
#pragma unroll
            for (int k = 0; k < 4; k++) x[k] = V[k][j];
        }
    }
}

__global__ void __launch_bounds__(128)
triangulate_kernel(fm3d_cam cam, const float* __restrict__ kp1, int n1, const float* __restrict__ kp2,
                   int n2, const int32_t* __restrict__ qidx, const int32_t* __restrict__ tidx, int n,
                   double* __restrict__ xyz_all, uint8_t* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int a = qidx ? qidx[i] : i, b = tidx ? tidx[i] : i;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (a < 0 || a >= n1 || b < 0 || b >= n2) {  // .at() would throw in the reference
        xyz_all[3 * i] = nan; xyz_all[3 * i + 1] = nan; xyz_all[3 * i + 2] = nan;
        mask[i] = 0;
        return;
    }
    double u1x, u1y, u2x, u2y;
    fm3d_undistort(cam, (double)kp1[2 * a], (double)kp1[2 * a + 1], u1x, u1y);
    fm3d_undistort(cam, (double)kp2[2 * b], (double)kp2[2 * b + 1], u2x, u2y);
    // rows x*P[2]-P[0], y*P[2]-P[1] for P1 = [I|0] and P2 = g12[0:3,:]
    double A[4][4];
    A[0][0] = -1; A[0][1] = 0; A[0][2] = u1x; A[0][3] = 0;
    A[1][0] = 0; A[1][1] = -1; A[1][2] = u1y; A[1][3] = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double p0 = k < 3 ? cam.R[k] : cam.t[0];
        const double p1 = k < 3 ? cam.R[3 + k] : cam.t[1];
        const double p2 = k < 3 ? cam.R[6 + k] : cam.t[2];
        A[2][k] = u2x * p2 - p0;
        A[3][k] = u2y * p2 - p1;
    }
    double hv[4];
    null_vector4(A, hv);
    const double X = hv[0] / hv[3], Y = hv[1] / hv[3], Z = hv[2] / hv[3];
    xyz_all[3 * i] = X; xyz_all[3 * i + 1] = Y; xyz_all[3 * i + 2] = Z;
    // drop iff Z/W < zmin || Z/W >= zmax (singlecameratriangulator.cpp:200); NaN counts as dropped
    const bool out = (Z < cam.zmin) || (Z >= cam.zmax) || (Z != Z);
    mask[i] = out ? 0 : 1;
}

// Order-preserving compaction by one CTA: running offset + block-wide ballot scan.
__global__ void __launch_bounds__(1024)
compact_kernel(const double* __restrict__ xyz_all, const uint8_t* __restrict__ mask, int n,
               double* __restrict__ xyz, int32_t* __restrict__ src_idx, int* __restrict__ ninl) {
    __shared__ int warp_cnt[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        const int i = start + tid;
        const bool keep = i < n && mask[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; w++) off += warp_cnt[w];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) {
            xyz[3 * off] = xyz_all[3 * i];
            xyz[3 * off + 1] = xyz_all[3 * i + 1];
            xyz[3 * off + 2] = xyz_all[3 * i + 2];
            if (src_idx) src_idx[off] = i;
        }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += warp_cnt[w];
            base_s += tot;
        }
        __syncthreads();
    }
    if (tid == 0) *ninl = base_s;
}

__global__ void undistort_kernel(fm3d_cam cam, const double* __restrict__ pts, int n,
                                 double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x, y;
    fm3d_undistort(cam, pts[2 * i], pts[2 * i + 1], x, y);
    out[2 * i] = x; out[2 * i + 1] = y;
}

}  // namespace

extern "C" {

int fm3d_triangulate_dev(fm3d_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                         const int32_t* qidx, const int32_t* tidx, int n, double* xyz_all,
                         uint8_t* mask, double* xyz, int32_t* src_idx, int* ninl_dev) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (!ctx->has_cam || !ctx->has_g12) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera / g12 not set");
    FM3D_CHECK_ARG(ctx, n >= 0 && n1 >= 0 && n2 >= 0 && (qidx == nullptr) == (tidx == nullptr));
    FM3D_CHECK_ARG(ctx, xyz_all && mask && xyz && ninl_dev && (n == 0 || (kp1 && kp2)));
    if (int rc = fm3d_bind(ctx)) return rc;
    if (n > 0) {
        triangulate_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->cam, kp1, n1, kp2, n2, qidx,
                                                                     tidx, n, xyz_all, mask);
        FM3D_LAUNCH_CHECK(ctx);
    }
    compact_kernel<<<1, 1024, 0, ctx->stream>>>(xyz_all, mask, n, xyz, src_idx, ninl_dev);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_triangulate(fm3d_ctx* ctx, const float* kp1, int n1, const float* kp2, int n2,
                     const int32_t* qidx, const int32_t* tidx, int n, double* xyz_all,
                     uint8_t* mask, double* xyz, int32_t* src_idx, int* ninl) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && n1 >= 0 && n2 >= 0 && ninl);
    FM3D_CHECK_ARG(ctx, n == 0 || (kp1 && kp2 && mask && xyz));
    if (int rc = fm3d_bind(ctx)) return rc;
    const size_t b_kp1 = sizeof(float) * 2 * (size_t)n1, b_kp2 = sizeof(float) * 2 * (size_t)n2;
    const size_t b_idx = qidx ? sizeof(int32_t) * (size_t)n : 0;
    const size_t b_xyz = sizeof(double) * 3 * (size_t)n;
    // layout of scratch slot 0: kp1 | kp2 | qidx | tidx | xyz_all | xyz | src | mask | ninl
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    size_t o_kp1 = 0, o_kp2 = o_kp1 + al(b_kp1), o_q = o_kp2 + al(b_kp2), o_t = o_q + al(b_idx);
    size_t o_all = o_t + al(b_idx), o_xyz = o_all + al(b_xyz), o_src = o_xyz + al(b_xyz);
    size_t o_mask = o_src + al(sizeof(int32_t) * (size_t)n), o_n = o_mask + al((size_t)n);
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, o_n + 256, (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_kp1, kp1, b_kp1)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_kp2, kp2, b_kp2)) return rc;
    if (qidx) {
        if (int rc = fm3d_h2d(ctx, d + o_q, qidx, b_idx)) return rc;
        if (int rc = fm3d_h2d(ctx, d + o_t, tidx, b_idx)) return rc;
    }
    int rc = fm3d_triangulate_dev(ctx, (const float*)(d + o_kp1), n1, (const float*)(d + o_kp2), n2,
                                  qidx ? (const int32_t*)(d + o_q) : nullptr,
                                  qidx ? (const int32_t*)(d + o_t) : nullptr, n, (double*)(d + o_all),
                                  (uint8_t*)(d + o_mask), (double*)(d + o_xyz), (int32_t*)(d + o_src),
                                  (int*)(d + o_n));
    if (rc) return rc;
    int h_n = 0;
    if (int rc2 = fm3d_d2h(ctx, &h_n, d + o_n, sizeof(int))) return rc2;
    if (xyz_all) if (int rc2 = fm3d_d2h(ctx, xyz_all, d + o_all, b_xyz)) return rc2;
    if (int rc2 = fm3d_d2h(ctx, mask, d + o_mask, (size_t)n)) return rc2;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (int rc2 = fm3d_d2h(ctx, xyz, d + o_xyz, sizeof(double) * 3 * (size_t)h_n)) return rc2;
    if (src_idx) if (int rc2 = fm3d_d2h(ctx, src_idx, d + o_src, sizeof(int32_t) * (size_t)h_n)) return rc2;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *ninl = h_n;
    return FM3D_OK;
}

int fm3d_undistort_points(fm3d_ctx* ctx, const double* pts, int n, double* out) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    if (!ctx->has_cam) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera not set");
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (pts && out)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    const size_t b = sizeof(double) * 2 * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 2 * b, (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, pts, b)) return rc;
    undistort_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->cam, (const double*)d, n, (double*)(d + b));
    FM3D_LAUNCH_CHECK(ctx);
    if (int rc = fm3d_d2h(ctx, out, d + b, b)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

}  // extern "C"
