// fm3d_normals.cu -- K5/K6/K7: the per-feature plane-normal search.
//
// Replaces NormalOptimizer::computeOptimizedNormals (Triangulator/normaloptimizer.cpp:321-452)
// and everything it drives: extractPixelsContour (singlecameratriangulator.cpp:341-397),
// optimize_pyramid / optimize (normaloptimizer.cpp:223-292), lmfit's lmmin, evaluateNormal
// (normaloptimizer.cpp:65-149) and its helpers get3dPointsFromImage1Pixels /
// updateImage1PixelsIntensity / projectPointsToImage2 / isInBoundingBox / isPixelGood
// (singlecameratriangulator.cpp:530-665) with getBilinearInterpPix32f (tools.cpp:129-142).
//
// Design (B200): persistent CTAs, one feature per CTA at a time, features pulled from an
// atomic queue (the number of LM evaluations varies 100..1200 per feature).
//   prologue   centre pixel (fp64), disc lattice in row-major order, 5-iteration undistort of
//              every disc pixel in fp64 -> ideal rays, stored as float2 offsets from the
//              centre ray in shared memory (rays do not depend on the normal: the reference
//              recomputes them in every evaluation, singlecameratriangulator.cpp:542)
//   per level  image-1 intensities sampled once (they cannot change within a level) and the
//              image-2 window around the projected feature staged in shared memory by one
//              TMA tensor-tile load (cp.async.bulk.tensor.2d, zero fill outside the image)
//   per pass   every thread walks its pixels: ray (R v precomputed once per pixel) ->
//              ray/plane intersection -> bounding-box test -> pose -> lens distortion -> K ->
//              x scale -> float cast -> 4-tap bilinear from the smem window (global fallback
//              with the reference's flat addressing when a tap leaves the window) -> residual.
//              A Jacobian pass evaluates f(x), f(x+h0 e0), f(x+h1 e1) for the same pixel and
//              accumulates six fp64 sums; a trial pass accumulates |f|^2.  Warp shuffles +
//              one smem stage reduce them; thread 0 advances the LM state machine
//              (fm3d_lm2.h) and publishes the next pass.
// The reference aborts a feature on the first bad pixel of any evaluation (D8); here the
// pass finishes and the OR of the per-pixel flags aborts it, which selects the same features.
#include "fm3d_normals_common.cuh"

using namespace fm3d_normals;

namespace {

struct LevelView {
    const uint8_t* img2;   // global, pitched
    int w, h, pitch;
    const uint8_t* win;    // shared window, pitch = ww
    int ww;
    int wx0, wy0;          // window origin in level pixels
    int lx_min, lx_cnt, ly_min, ly_cnt;  // taps (x0,y0) whose 2x2 footprint is inside window AND image
};

// getBilinearInterpPix32f on image 2 of the current level: 2x2 taps from the staged window when
// they are all inside it (and inside the image), else global memory with flat addressing.
__device__ __forceinline__ float sample_img2(const LevelView& L, float x, float y) {
    const float fx0 = floorf(x), fy0 = floorf(y);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const int lx = x0 - L.wx0, ly = y0 - L.wy0;
    float b00, b01, b10, b11;
    if ((unsigned)(lx - L.lx_min) < (unsigned)L.lx_cnt && (unsigned)(ly - L.ly_min) < (unsigned)L.ly_cnt) {
        const uint8_t* p = L.win + ly * L.ww + lx;
        b00 = fm3d_u8f(p[0]); b10 = fm3d_u8f(p[1]);
        b01 = fm3d_u8f(p[L.ww]); b11 = fm3d_u8f(p[L.ww + 1]);
    } else {
        b00 = fm3d_u8f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0, y0));
        b01 = fm3d_u8f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0, y0 + 1));
        b10 = fm3d_u8f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0 + 1, y0));
        b11 = fm3d_u8f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0 + 1, y0 + 1));
    }
    return fm3d_lerp4(b00, b01, b10, b11, __fsub_rn(x, fx0), __fsub_rn(y, fy0));
}

// Ray offsets from the centre ray, 8 bytes per pixel.  fp64 geometry stores them as 32-bit fixed
// point (resolution <= 2^-31 of the disc's angular extent, ~1e-10: two orders finer than float),
// fp32 geometry stores floats.
template <typename G> struct RayStore;
template <> struct RayStore<double> {
    typedef int2 type;
    static __device__ __forceinline__ int2 pack(double dx, double dy, double scale) {
        return make_int2(__double2int_rn(dx * scale), __double2int_rn(dy * scale));
    }
    static __device__ __forceinline__ void unpack(int2 v, double inv_scale, double cx, double cy, double& x, double& y) {
        x = fma((double)v.x, inv_scale, cx);
        y = fma((double)v.y, inv_scale, cy);
    }
};
template <> struct RayStore<float> {
    typedef float2 type;
    static __device__ __forceinline__ float2 pack(double dx, double dy, double) { return make_float2((float)dx, (float)dy); }
    static __device__ __forceinline__ void unpack(float2 v, float, float cx, float cy, float& x, float& y) {
        x = cx + v.x;
        y = cy + v.y;
    }
};

template <typename G>
struct CamG {
    G fx, fy, cx, cy, k1, k2, p1, p2, k3;
    G R[9], t[3];
    G cmax;
    G scale, inv_scale_cols, inv_scale_rows;
};

// One evaluation of one pixel: returns the unweighted float residual I1 - I2 and ORs flags.
template <typename G>
__device__ __forceinline__ float eval_pixel(const CamG<G>& C, const LevelView& L, G vx, G vy, G a, G b,
                                            G c, G nx, G ny, G nz, G mnum, float I1, unsigned& flags) {
    // projectPointToPlane (singlecameratriangulator.cpp:421-470): k = (n.P)/(n.v), X = k v
    const G den = nx * vx + ny * vy + nz;
    const G k = mnum / den;
    const G X = k * vx, Y = k * vy;
    if ((X != X) || (Y != Y) || (k != k)) flags |= FLAG_NAN;
    // isInBoundingBox (:646-655)
    if (!((X > -C.cmax && X < C.cmax) && (Y > -C.cmax && Y < C.cmax) && (k > G(0) && k < C.cmax))) flags |= FLAG_BBOX;
    // cv::projectPoints with (r2,t2) = g12 (:591-602): R(k v) + t = k (R v) + t
    const G x2 = k * a + C.t[0], y2 = k * b + C.t[1], z2 = k * c + C.t[2];
    const G iz = z2 != G(0) ? G(1) / z2 : G(1);
    G u, v;
    fm3d_distort_K<G>(x2 * iz, y2 * iz, C.k1, C.k2, C.p1, C.p2, C.k3, C.fx, C.fy, C.cx, C.cy, u, v);
    // isPixelGood (:657-665)
    if ((u < G(0)) || (u > C.inv_scale_cols) || (v < G(0)) || (v > C.inv_scale_rows) || (u != u) || (v != v))
        flags |= FLAG_PIX;
    const float su = (float)(C.scale * u), sv = (float)(C.scale * v);
    const float I2 = sample_img2(L, su, sv);
    return __fsub_rn(I1, I2);
}
template <typename G, bool RAYS_SMEM>
__global__ void __launch_bounds__(NT_MAX, 1)
normals_kernel(const __grid_constant__ NormalsArgs A) {
    extern __shared__ __align__(128) uint8_t smem[];
    typedef typename RayStore<G>::type ray_t;
    uint8_t* win = smem;
    ray_t* rays;
    float* i1;
    uint8_t* tail;
    if (RAYS_SMEM) {
        rays = reinterpret_cast<ray_t*>(smem + WIN_BYTES);
        i1 = reinterpret_cast<float*>(rays + A.mcap);
        tail = reinterpret_cast<uint8_t*>(i1 + A.mcap);
    } else {
        rays = reinterpret_cast<ray_t*>(A.rays_g) + (size_t)blockIdx.x * A.mcap;
        i1 = A.i1_g + (size_t)blockIdx.x * A.mcap;
        tail = smem + WIN_BYTES;
    }
    tail = reinterpret_cast<uint8_t*>(((uintptr_t)tail + 15) & ~(uintptr_t)15);
    RowTable* rows = reinterpret_cast<RowTable*>(tail);
    tail += (sizeof(RowTable) + 15) & ~(size_t)15;
    double* red = reinterpret_cast<double*>(tail);        // [16 warps][6]
    tail += sizeof(double) * 16 * 6;
    unsigned* wflags = reinterpret_cast<unsigned*>(tail);  // [16 warps]
    tail += sizeof(unsigned) * 16;
    FeatureShared* S = reinterpret_cast<FeatureShared*>(tail);
    tail += (sizeof(FeatureShared) + 15) & ~(size_t)15;
    PassParams<G>* PP = reinterpret_cast<PassParams<G>*>(tail);
    tail += (sizeof(PassParams<G>) + 15) & ~(size_t)15;
    uint64_t* bar = reinterpret_cast<uint64_t*>(tail);

    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    const fm3d_cam& cam = A.cam;
    const int r = A.r, W = A.pyr.lv[0].w, H = A.pyr.lv[0].h, levels = A.pyr.levels;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        S->tma_phase = 0;
    }
    __syncthreads();

    CamG<G> C;
    C.fx = (G)cam.fx; C.fy = (G)cam.fy; C.cx = (G)cam.cx; C.cy = (G)cam.cy;
    C.k1 = (G)cam.k1; C.k2 = (G)cam.k2; C.p1 = (G)cam.p1; C.p2 = (G)cam.p2; C.k3 = (G)cam.k3;
#pragma unroll
    for (int i = 0; i < 9; i++) C.R[i] = (G)cam.R[i];
#pragma unroll
    for (int i = 0; i < 3; i++) C.t[i] = (G)cam.t[i];
    C.cmax = (G)(int)(2 * cam.zmax);  // int cMax = 2*z_threshold_max_ (:648)

    for (;;) {
        // ------------------------------------------------------------ fetch a feature
        if (tid == 0) S->feature = atomicAdd(A.work_counter, 1);
        __syncthreads();
        const int f = S->feature;
        if (f >= A.n) break;
        const double Px = A.xyz[3 * f], Py = A.xyz[3 * f + 1], Pz = A.xyz[3 * f + 2];

        // ------------------------------------------------------------ prologue: disc lattice
        double cu, cv;
        fm3d_project(cam, Px, Py, Pz, cu, cv);  // extractPixelsContour(Vec3d) (:376-397)
        for (int jr = tid; jr < 2 * r + 1; jr += NT) {
            const int j = jr - r;
            const double py = cv + (double)j;
            int cnt = 0, lo = 0;
            if (!(py < 0 || py >= (double)H) && cu == cu && cv == cv) {
                const int hw = (int)floor(sqrt((double)(r * r - j * j)));
                int ilo = -hw, ihi = hw;
                // keep iff !(px < 0 || px >= W) with px = cu + i evaluated exactly as the reference does
                if (cu + (double)ilo < 0) {
                    int g = (int)ceil(-cu);
                    ilo = g < -hw ? -hw : (g > hw + 1 ? hw + 1 : g);
                    while (ilo <= hw && (cu + (double)ilo) < 0) ilo++;
                    while (ilo - 1 >= -hw && !((cu + (double)(ilo - 1)) < 0)) ilo--;
                }
                if (cu + (double)ihi >= (double)W) {
                    int g = (int)ceil((double)W - cu) - 1;
                    ihi = g > hw ? hw : (g < -hw - 1 ? -hw - 1 : g);
                    while (ihi >= -hw && (cu + (double)ihi) >= (double)W) ihi--;
                    while (ihi + 1 <= hw && !((cu + (double)(ihi + 1)) >= (double)W)) ihi++;
                }
                cnt = ihi - ilo + 1;
                if (cnt < 0) cnt = 0;
                lo = ilo;
            }
            rows->start[jr + 1] = cnt;  // counts, prefix-summed below
            rows->ilo[jr] = (short)lo;
            rows->jrow[jr] = (short)j;
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            rows->start[0] = 0;
            for (int jr = 0; jr < 2 * r + 1; jr++) { acc += rows->start[jr + 1]; rows->start[jr + 1] = acc; }
            rows->nrows = 2 * r + 1;
            S->m = acc;
            S->status = acc > 0 ? FM3D_FEAT_OK : FM3D_FEAT_NO_PIXELS;
            // a non-finite centre passes every `p < 0 || p >= size` test of the reference's lattice loop,
            // the feature then dies in its first evaluation with a NaN plane point
            if (cu != cu || cv != cv) S->status = FM3D_FEAT_ABORT_NAN;
            S->npenalty = 0;
            S->P[0] = Px; S->P[1] = Py; S->P[2] = Pz;
            const double nrm = sqrt(Px * Px + Py * Py + Pz * Pz);
            S->normal[0] = Px / nrm; S->normal[1] = Py / nrm; S->normal[2] = Pz / nrm;  // (:343)
            if (A.nfev && A.mode == 0) for (int l = 0; l <= levels; l++) A.nfev[(size_t)f * (levels + 1) + l] = 0;
        }
        __syncthreads();
        const int m = S->m;
        if (A.m_out && tid == 0) A.m_out[f] = m;

        // centre ray and ideal rays of all disc pixels (normal-independent)
        double vcx, vcy;
        fm3d_undistort(cam, cu, cv, vcx, vcy);
        // fixed-point scale: offsets are bounded by ~(r+2)/f (x2 for lens distortion)
        const double ray_bound = 2.0 * (double)(r + 2) / fmin(fabs(cam.fx), fabs(cam.fy));
        const double ray_scale = ldexp(1.0, 30 - ilogb(ray_bound));
        const G ray_inv_scale = (G)(1.0 / ray_scale);
        if (m > 0) {
            int row = 0;
            for (int idx = tid; idx < m; idx += NT) {
                while (idx >= rows->start[row + 1]) row++;
                const double px = cu + (double)(rows->ilo[row] + (idx - rows->start[row]));
                const double py = cv + (double)rows->jrow[row];
                double vx, vy;
                fm3d_undistort(cam, px, py, vx, vy);
                rays[idx] = RayStore<G>::pack(vx - vcx, vy - vcy, ray_scale);
            }
        }
        const G gvcx = (G)vcx, gvcy = (G)vcy;

        // ------------------------------------------------------------ coarse-to-fine LM
        const int lvl_hi = A.mode == 0 ? levels : A.eval_level;
        const int lvl_lo = A.mode == 0 ? 0 : A.eval_level;
        bool alive = m > 0;
        for (int lvl = lvl_hi; lvl >= lvl_lo && alive; lvl--) {
            const fm3d_level lv = A.pyr.lv[lvl];
            const double scale = 1.0 / (double)(1 << lvl);      // actual_scale_ (:226-241)
            const double inv_scale = 1.0 / scale;
            const uint8_t* img1 = A.pyr.base[0] + lv.off;
            LevelView L;
            L.img2 = A.pyr.base[1] + lv.off;
            L.w = lv.w; L.h = lv.h; L.pitch = lv.pitch;
            L.win = win; L.ww = A.win_w[lvl];
            const int wh = A.win_h[lvl];
            C.scale = (G)scale;
            C.inv_scale_cols = (G)(inv_scale * lv.w);
            C.inv_scale_rows = (G)(inv_scale * lv.h);

            // window origin: centred on the projection of P into image 2 (P lies on every candidate plane)
            if (tid == 0) {
                const double X2 = cam.R[0] * Px + cam.R[1] * Py + cam.R[2] * Pz + cam.t[0];
                const double Y2 = cam.R[3] * Px + cam.R[4] * Py + cam.R[5] * Pz + cam.t[1];
                const double Z2 = cam.R[6] * Px + cam.R[7] * Py + cam.R[8] * Pz + cam.t[2];
                double u2, v2;
                fm3d_project(cam, X2, Y2, Z2, u2, v2);
                double wxc = floor(scale * u2), wyc = floor(scale * v2);
                if (!(wxc > -1e6 && wxc < 1e6)) wxc = 0;
                if (!(wyc > -1e6 && wyc < 1e6)) wyc = 0;
                // TMA needs the box start address 16-byte aligned: x origin is a multiple of 16 pixels
                S->wx0 = (((int)wxc - L.ww / 2 + 8) >> 4) << 4;
                S->wy0 = (int)wyc - wh / 2;
                S->level = lvl;
            }
            __syncthreads();  // also: everybody is done with the previous level's window
            L.wx0 = S->wx0; L.wy0 = S->wy0;
            bool staged = false;
            if (L.ww > 0 && wh > 0) {
                if (A.use_tma) {
                    const uint32_t parity = (uint32_t)S->tma_phase;
                    if (tid == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(bar, (uint32_t)(L.ww * wh));
                        tma_load_2d(win, &A.tmap[lvl], L.wx0, L.wy0, bar);
                    }
                    bool ok = false;
                    for (int spin = 0; spin < (1 << 22); spin++) {
                        if (mbar_try_wait(bar, parity)) { ok = true; break; }
                    }
                    staged = __syncthreads_and(ok ? 1 : 0) != 0;
                    if (tid == 0) {
                        S->tma_phase ^= 1;
                        if (!staged) atomicExch(A.error_flag, 1);
                    }
                }
                if (!staged && !A.use_tma) {
                    for (int i = tid; i < L.ww * wh; i += NT) {
                        const int yy = i / L.ww, xx = i - yy * L.ww;
                        const int gx = L.wx0 + xx, gy = L.wy0 + yy;
                        win[i] = (gx >= 0 && gx < lv.w && gy >= 0 && gy < lv.h) ? L.img2[(size_t)gy * lv.pitch + gx] : 0;
                    }
                    staged = true;
                }
            }
            // taps (x0,y0),(x0+1,y0+1) must lie inside the window and inside the image
            {
                int x_lo = max(L.wx0, 0), x_hi = min(L.wx0 + L.ww, lv.w) - 1;  // x0 in [x_lo, x_hi)
                int y_lo = max(L.wy0, 0), y_hi = min(L.wy0 + wh, lv.h) - 1;
                L.lx_min = x_lo - L.wx0; L.lx_cnt = staged ? max(x_hi - x_lo, 0) : 0;
                L.ly_min = y_lo - L.wy0; L.ly_cnt = staged ? max(y_hi - y_lo, 0) : 0;
            }

            // image-1 intensities of the level (updateImage1PixelsIntensity, :576-589)
            unsigned lvl_flags = 0;
            {
                int row = 0;
                for (int idx = tid; idx < m; idx += NT) {
                    while (idx >= rows->start[row + 1]) row++;
                    const double px = cu + (double)(rows->ilo[row] + (idx - rows->start[row]));
                    const double py = cv + (double)rows->jrow[row];
                    if (!fm3d_pixel_good(px, py, inv_scale, lv.w, lv.h)) lvl_flags |= FLAG_PIX;
                    i1[idx] = fm3d_bilinear_global(img1, lv.w, lv.h, lv.pitch, (float)(scale * px), (float)(scale * py));
                }
            }

            // thread 0: start the LM of this level (optimize(), :247-292)
            if (tid == 0) {
                double phi, theta;
                if (A.mode == 0) {
                    const double* nv = S->normal;
                    theta = atan2(nv[2], sqrt(nv[0] * nv[0] + nv[1] * nv[1]));  // car2sph (tools.cpp:767-771)
                    phi = atan2(nv[1], nv[0]);
                    PP->cmd = fm3d_lm2_init(&S->lm, phi, theta, A.eps_lmmin, A.patience);
                } else {
                    phi = A.phi_theta[2 * f]; theta = A.phi_theta[2 * f + 1];
                    S->lm.xt[0] = phi; S->lm.xt[1] = theta;
                    S->lm.nfev = 0;
                    PP->cmd = FM3D_LM_CMD_TRIAL;
                }
            }
            __syncthreads();

            // -------------------------------------------------------- pass loop
            for (;;) {
                // thread 0 publishes the evaluation points of the pass
                if (tid == 0) {
                    const int cmd = PP->cmd;
                    const fm3d_lm2& lm = S->lm;
                    double ph[3], th[3];
                    int ne;
                    if (cmd == FM3D_LM_CMD_JAC) {
                        ne = 3;
                        ph[0] = lm.x[0]; th[0] = lm.x[1];
                        ph[1] = lm.x[0] + lm.h[0]; th[1] = lm.x[1];
                        ph[2] = lm.x[0]; th[2] = lm.x[1] + lm.h[1];
                    } else {
                        ne = 1;
                        ph[0] = lm.xt[0]; th[0] = lm.xt[1];
                    }
                    bool nan_normal = false;
                    for (int e = 0; e < ne; e++) {
                        // sph2car (tools.cpp:772-777)
                        const double n0 = cos(th[e]) * cos(ph[e]), n1 = cos(th[e]) * sin(ph[e]), n2 = sin(th[e]);
                        if (n0 != n0 || n1 != n1 || n2 != n2) nan_normal = true;  // (:81-85)
                        PP->nx[e] = (G)n0; PP->ny[e] = (G)n1; PP->nz[e] = (G)n2;
                        PP->mnum[e] = (G)(n0 * Px + n1 * Py + n2 * Pz);
                        int entered;
                        S->w[e] = penalty_weight(ph[e], th[e], A.penalty_mode, entered);
                        // count the evaluations the reference would make: f(x) of a Jacobian pass is a
                        // re-evaluation except in the very first pass of the level
                        if (entered && !(cmd == FM3D_LM_CMD_JAC && e == 0 && !lm.first)) S->npenalty++;
                    }
                    PP->ne = ne;
                    // ne == 0 tells everybody to abort; PP->cmd must not change here (the other
                    // threads may still be reading it at the bottom of the previous iteration)
                    if (nan_normal) { S->status = FM3D_FEAT_ABORT_NAN; PP->ne = 0; }
                }
                __syncthreads();
                const int ne = PP->ne;
                if (ne == 0) { alive = false; break; }

                unsigned flags = lvl_flags;
                lvl_flags = 0;
                double acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0;
                if (ne == 3) {
                    const G n0x = PP->nx[0], n0y = PP->ny[0], n0z = PP->nz[0], m0 = PP->mnum[0];
                    const G n1x = PP->nx[1], n1y = PP->ny[1], n1z = PP->nz[1], m1 = PP->mnum[1];
                    const G n2x = PP->nx[2], n2y = PP->ny[2], n2z = PP->nz[2], m2 = PP->mnum[2];
                    for (int idx = tid; idx < m; idx += NT) {
                        const ray_t dr = rays[idx];
                        const float I1 = i1[idx];
                        G vx, vy;
                        RayStore<G>::unpack(dr, ray_inv_scale, gvcx, gvcy, vx, vy);
                        const G a = C.R[0] * vx + C.R[1] * vy + C.R[2];
                        const G b = C.R[3] * vx + C.R[4] * vy + C.R[5];
                        const G c = C.R[6] * vx + C.R[7] * vy + C.R[8];
                        unsigned f1 = 0, f2 = 0;
                        const float d0 = eval_pixel<G>(C, L, vx, vy, a, b, c, n0x, n0y, n0z, m0, I1, flags);
                        const float d1 = eval_pixel<G>(C, L, vx, vy, a, b, c, n1x, n1y, n1z, m1, I1, f1);
                        const float d2 = eval_pixel<G>(C, L, vx, vy, a, b, c, n2x, n2y, n2z, m2, I1, f2);
                        flags |= (f1 << 3) | (f2 << 6);
                        const double D0 = (double)d0;
                        const double e1 = (double)__fsub_rn(d1, d0), e2 = (double)__fsub_rn(d2, d0);
                        acc0 = fma(D0, D0, acc0);
                        acc1 = fma(e1, e1, acc1);
                        acc2 = fma(e1, e2, acc2);
                        acc3 = fma(e2, e2, acc3);
                        acc4 = fma(e1, D0, acc4);
                        acc5 = fma(e2, D0, acc5);
                    }
                } else {
                    const G n0x = PP->nx[0], n0y = PP->ny[0], n0z = PP->nz[0], m0 = PP->mnum[0];
                    for (int idx = tid; idx < m; idx += NT) {
                        const ray_t dr = rays[idx];
                        const float I1 = i1[idx];
                        G vx, vy;
                        RayStore<G>::unpack(dr, ray_inv_scale, gvcx, gvcy, vx, vy);
                        const G a = C.R[0] * vx + C.R[1] * vy + C.R[2];
                        const G b = C.R[3] * vx + C.R[4] * vy + C.R[5];
                        const G c = C.R[6] * vx + C.R[7] * vy + C.R[8];
                        const float d0 = eval_pixel<G>(C, L, vx, vy, a, b, c, n0x, n0y, n0z, m0, I1, flags);
                        const double D0 = (double)d0;
                        acc0 = fma(D0, D0, acc0);
                    }
                }
                acc0 = warp_sum(acc0);
                if (ne == 3) {
                    acc1 = warp_sum(acc1); acc2 = warp_sum(acc2); acc3 = warp_sum(acc3);
                    acc4 = warp_sum(acc4); acc5 = warp_sum(acc5);
                }
                flags = __reduce_or_sync(0xffffffffu, flags);
                if (lane == 0) {
                    double* rw = red + wid * 6;
                    rw[0] = acc0; rw[1] = acc1; rw[2] = acc2; rw[3] = acc3; rw[4] = acc4; rw[5] = acc5;
                    wflags[wid] = flags;
                }
                __syncthreads();

                if (tid == 0) {
                    unsigned any_flags = 0;
                    for (int w = 0; w < NW; w++) any_flags |= wflags[w];
                    if (any_flags) {
                        // flag groups of f(x), f(x+h0 e0), f(x+h1 e1): the reference evaluates them in this
                        // order and stops at the first failing one; inside one evaluation all bounding-box /
                        // NaN tests precede the pixel tests
                        int g = any_flags & 7;
                        if (!g) g = (any_flags >> 3) & 7;
                        if (!g) g = (any_flags >> 6) & 7;
                        S->status = (g & FLAG_NAN) ? FM3D_FEAT_ABORT_NAN
                                  : (g & FLAG_BBOX) ? FM3D_FEAT_ABORT_BBOX : FM3D_FEAT_ABORT_PIXEL;
                        PP->cmd = 0;
                    } else {
                        double s[6] = {0, 0, 0, 0, 0, 0};
                        for (int w = 0; w < NW; w++)
                            for (int k = 0; k < 6; k++) s[k] += red[w * 6 + k];
                        fm3d_lm2& lm = S->lm;
                        if (A.mode != 0) {
                            A.cost[f] = S->w[0] * S->w[0] * s[0];
                            PP->cmd = FM3D_LM_CMD_DONE;
                        } else if (PP->cmd == FM3D_LM_CMD_JAC) {
                            // f_e = w_e d_e;  J_j = (f_j - f_0)/h_j = (w_j e_j + (w_j - w_0) d_0)/h_j
                            const double w0 = S->w[0], w1 = S->w[1], w2 = S->w[2];
                            const double a1 = w1 - w0, a2 = w2 - w0;
                            const double A00 = s[0], E11 = s[1], E12 = s[2], E22 = s[3], E1d = s[4], E2d = s[5];
                            const double ih0 = 1.0 / lm.h[0], ih1 = 1.0 / lm.h[1];
                            const double ff = w0 * w0 * A00;
                            const double S00 = (w1 * w1 * E11 + 2 * w1 * a1 * E1d + a1 * a1 * A00) * ih0 * ih0;
                            const double S11 = (w2 * w2 * E22 + 2 * w2 * a2 * E2d + a2 * a2 * A00) * ih1 * ih1;
                            const double S01 = (w1 * w2 * E12 + w1 * a2 * E1d + w2 * a1 * E2d + a1 * a2 * A00) * ih0 * ih1;
                            const double g0 = w0 * (w1 * E1d + a1 * A00) * ih0;
                            const double g1 = w0 * (w2 * E2d + a2 * A00) * ih1;
                            PP->cmd = fm3d_lm2_after_jacobian(&lm, ff, S00, S01, S11, g0, g1);
                        } else {
                            PP->cmd = fm3d_lm2_after_trial(&lm, S->w[0] * S->w[0] * s[0]);
                        }
                    }
                }
                __syncthreads();
                const int cmd = PP->cmd;
                if (cmd == 0) { alive = false; break; }
                if (cmd == FM3D_LM_CMD_DONE) break;
            }

            if (tid == 0 && alive && A.mode == 0) {
                const fm3d_lm2& lm = S->lm;
                // sph2car of the solution (:289)
                S->normal[0] = cos(lm.x[1]) * cos(lm.x[0]);
                S->normal[1] = cos(lm.x[1]) * sin(lm.x[0]);
                S->normal[2] = sin(lm.x[1]);
                if (A.nfev) A.nfev[(size_t)f * (levels + 1) + lvl] = lm.nfev;
                if (A.cost) A.cost[f] = lm.ff;
            }
            if (tid == 0 && !alive && A.mode == 0 && A.nfev) A.nfev[(size_t)f * (levels + 1) + lvl] = S->lm.nfev;
            __syncthreads();
        }

        // ------------------------------------------------------------ epilogue
        if (tid == 0) {
            const int st = S->status;
            A.status[f] = st;
            if (A.mode == 0) {
                if (st == FM3D_FEAT_OK) {
                    A.normals[3 * f] = S->normal[0]; A.normals[3 * f + 1] = S->normal[1]; A.normals[3 * f + 2] = S->normal[2];
                } else {
                    const double nrm = sqrt(Px * Px + Py * Py + Pz * Pz);
                    A.normals[3 * f] = Px / nrm; A.normals[3 * f + 1] = Py / nrm; A.normals[3 * f + 2] = Pz / nrm;
                    if (A.cost) A.cost[f] = __longlong_as_double(0x7ff8000000000000LL);
                }
                if (A.npenalty) A.npenalty[f] = S->npenalty;
            } else if (st != FM3D_FEAT_OK) {
                A.cost[f] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        __syncthreads();
    }
}

size_t tail_bytes(bool f32) {
    size_t t = 16 + ((sizeof(RowTable) + 15) & ~(size_t)15) + sizeof(double) * 16 * 6 + sizeof(unsigned) * 16 +
               ((sizeof(FeatureShared) + 15) & ~(size_t)15);
    t += f32 ? ((sizeof(PassParams<float>) + 15) & ~(size_t)15) : ((sizeof(PassParams<double>) + 15) & ~(size_t)15);
    return t + 16;
}
__global__ void frames_kernel(const double* __restrict__ xyz, const double* __restrict__ normals, int n,
                              double gx, double gy, double gz, double* __restrict__ frames) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // computeFeaturesFrames (normaloptimizer.cpp:454-505): z = n, x = g x z, y = z x x, normalised
    const double zx = normals[3 * i], zy = normals[3 * i + 1], zz = normals[3 * i + 2];
    double xx = gy * zz - gz * zy, xy = gz * zx - gx * zz, xz = gx * zy - gy * zx;
    double yx = zy * xz - zz * xy, yy = zz * xx - zx * xz, yz = zx * xy - zy * xx;
    const double nx = sqrt(xx * xx + xy * xy + xz * xz), ny = sqrt(yx * yx + yy * yy + yz * yz);
    xx /= nx; xy /= nx; xz /= nx;
    yx /= ny; yy /= ny; yz /= ny;
    double* F = frames + 16 * (size_t)i;
    F[0] = xx; F[1] = yx; F[2] = zx; F[3] = xyz[3 * i];
    F[4] = xy; F[5] = yy; F[6] = zy; F[7] = xyz[3 * i + 1];
    F[8] = xz; F[9] = yz; F[10] = zz; F[11] = xyz[3 * i + 2];
    F[12] = 0; F[13] = 0; F[14] = 0; F[15] = 1;
}

int run_normals(fm3d_ctx* ctx, NormalsArgs& A) {
    if (!ctx->has_cam || !ctx->has_g12) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera / g12 not set");
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    if (A.r < 0 || A.r > MAX_RAY) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "pixels_ray must be in [0,%d]", MAX_RAY);
    if (A.n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    // default: the fast kernel for the optimisation; cost evaluations (parity probes) stay on the
    // evaluation-by-evaluation kernel unless normals_fast >= 2
    if (ctx->opt_normals_fast && (A.mode == 0 || ctx->opt_normals_fast >= 2)) return run_normals_fast(ctx, A);
    if (ctx->opt_normals_cost != FM3D_COST_SSD)
        return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "normals_cost = NCC runs in the fast kernel only (normals_fast >= 1; fm3d_evaluate_normals: normals_fast = 2)");
    A.cam = ctx->cam;
    A.pyr = ctx->pyr;
    A.patience = ctx->opt_lm_patience;
    A.use_tma = ctx->opt_normals_tma;
    A.mcap = (disc_capacity(A.r) + 31) & ~31;
    const bool f32 = ctx->opt_geometry_f32 != 0;
    int nt = ctx->opt_normals_threads;
    nt = nt < 64 ? 64 : (nt > NT_MAX ? NT_MAX : (nt & ~31));

    // window per level: the warp of the disc is close to a similarity, so 1.3x the scaled radius
    for (int l = 0; l <= A.pyr.levels; l++) {
        const double rs = (double)A.r / (double)(1 << l);
        int half = (int)ceil(1.3 * rs) + 8;
        int ww = (2 * half + 16 + 15) & ~15, wh = 2 * half;   // +16: slack for the 16-pixel origin alignment
        if (ww > WIN_MAX_W) ww = WIN_MAX_W;
        if (wh > WIN_MAX_H) wh = WIN_MAX_H;
        A.win_w[l] = ww; A.win_h[l] = wh;
        if (A.use_tma) {
            const fm3d_level& lv = A.pyr.lv[l];
            if (int rc = fm3d_encode_tmap_2d_u8(ctx, &A.tmap[l], A.pyr.base[1] + lv.off, lv.w, lv.h, lv.pitch, ww, wh))
                return rc;
        }
    }
    const size_t smem_max = ctx->prop.sharedMemPerBlockOptin;
    const size_t per_px = sizeof(float2) + sizeof(float);
    size_t smem_in = WIN_BYTES + per_px * (size_t)A.mcap + tail_bytes(f32) + 128;
    const bool rays_smem = smem_in <= smem_max;
    const size_t smem = rays_smem ? smem_in : (size_t)WIN_BYTES + tail_bytes(f32) + 128;

    auto launch = [&](auto kernel) -> int {
        FM3D_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        FM3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nt, smem));
        if (occ < 1) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "normals kernel does not fit (smem %zu)", smem);
        int grid = ctx->prop.multiProcessorCount * occ;
        if (grid > A.n) grid = A.n;
        int* ctrl = nullptr;  // [0] work counter, [1] error flag
        if (int rc = fm3d_scratch(ctx, 1, 256, (void**)&ctrl)) return rc;
        ctx->n_copy++;
        FM3D_CUDA(ctx, cudaMemsetAsync(ctrl, 0, 256, ctx->stream));
        A.work_counter = ctrl;
        A.error_flag = ctrl + 1;
        A.rays_g = nullptr; A.i1_g = nullptr;
        if (!rays_smem) {
            char* g = nullptr;
            if (int rc = fm3d_scratch(ctx, 2, per_px * (size_t)A.mcap * grid, (void**)&g)) return rc;
            A.rays_g = (float2*)g;
            A.i1_g = (float*)(g + sizeof(float2) * (size_t)A.mcap * grid);
        }
        kernel<<<grid, nt, smem, ctx->stream>>>(A);
        FM3D_LAUNCH_CHECK(ctx);
        return FM3D_OK;
    };
    if (f32) return rays_smem ? launch(normals_kernel<float, true>) : launch(normals_kernel<float, false>);
    return rays_smem ? launch(normals_kernel<double, true>) : launch(normals_kernel<double, false>);
}

}  // namespace

extern "C" {

int fm3d_get_normals_stats(fm3d_ctx* ctx, int64_t out[16]) {
    if (!ctx || !out) return FM3D_ERR_INVALID_ARG;
    for (int k = 0; k < 16; k++) out[k] = 0;
    if (!ctx->scratch[1]) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    FM3D_CUDA(ctx, cudaMemcpyAsync(out, (const char*)ctx->scratch[1] + 64, 16 * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_optimize_normals_dev(fm3d_ctx* ctx, const double* xyz, int n, int pixels_ray,
                              double epsilon_lmmin, int penalty_mode, double* normals,
                              int32_t* status, int32_t* nfev, int32_t* npenalty, double* cost) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (xyz && normals && status)));
    FM3D_CHECK_ARG(ctx, penalty_mode >= 0 && penalty_mode <= 2);
    NormalsArgs A{};
    A.xyz = xyz; A.n = n; A.r = pixels_ray; A.eps_lmmin = epsilon_lmmin; A.penalty_mode = penalty_mode;
    A.mode = 0;
    A.normals = normals; A.status = status; A.nfev = nfev; A.npenalty = npenalty; A.cost = cost;
    ctx->normals_flag_pending = true;       // looked at by fm3d_sync (or by the host-buffer entry point below)
    return run_normals(ctx, A);
}

int fm3d_optimize_normals(fm3d_ctx* ctx, const double* xyz, int n, int pixels_ray,
                          double epsilon_lmmin, int penalty_mode, double* normals,
                          int32_t* status, int32_t* nfev, int32_t* npenalty, double* cost) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (xyz && normals && status)));
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    const int L1 = ctx->pyr.levels + 1;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b3 = sizeof(double) * 3 * (size_t)n, bi = sizeof(int32_t) * (size_t)n;
    size_t o_xyz = 0, o_nrm = o_xyz + al(b3), o_st = o_nrm + al(b3), o_nf = o_st + al(bi);
    size_t o_np = o_nf + al(bi * L1), o_c = o_np + al(bi), o_end = o_c + al(sizeof(double) * (size_t)n);
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, o_end, (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_xyz, xyz, b3)) return rc;
    int rc = fm3d_optimize_normals_dev(ctx, (const double*)(d + o_xyz), n, pixels_ray, epsilon_lmmin, penalty_mode,
                                       (double*)(d + o_nrm), (int32_t*)(d + o_st), (int32_t*)(d + o_nf),
                                       (int32_t*)(d + o_np), (double*)(d + o_c));
    if (rc) return rc;
    if (int r2 = fm3d_d2h(ctx, normals, d + o_nrm, b3)) return r2;
    if (int r2 = fm3d_d2h(ctx, status, d + o_st, bi)) return r2;
    if (nfev) if (int r2 = fm3d_d2h(ctx, nfev, d + o_nf, bi * L1)) return r2;
    if (npenalty) if (int r2 = fm3d_d2h(ctx, npenalty, d + o_np, bi)) return r2;
    if (cost) if (int r2 = fm3d_d2h(ctx, cost, d + o_c, sizeof(double) * (size_t)n)) return r2;
    int flags[2] = {0, 0};
    if (int r2 = fm3d_d2h(ctx, flags, ctx->scratch[1], sizeof(flags))) return r2;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->normals_flag_pending = false;
    if (flags[1]) return fm3d_fail(ctx, FM3D_ERR_CUDA, "normal optimiser: a TMA window load timed out (results used the global-memory path)");
    return FM3D_OK;
}

int fm3d_evaluate_normals(fm3d_ctx* ctx, const double* xyz, const double* normals_phi_theta, int n,
                          int pixels_ray, int level, int penalty_mode, double* cost, int32_t* m,
                          int32_t* status) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (xyz && normals_phi_theta && cost && m && status)));
    FM3D_CHECK_ARG(ctx, penalty_mode >= 0 && penalty_mode <= 2);
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    FM3D_CHECK_ARG(ctx, level >= 0 && level <= ctx->pyr.levels);
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b3 = sizeof(double) * 3 * (size_t)n, b2 = sizeof(double) * 2 * (size_t)n, bi = sizeof(int32_t) * (size_t)n;
    size_t o_xyz = 0, o_pt = o_xyz + al(b3), o_c = o_pt + al(b2), o_m = o_c + al(sizeof(double) * (size_t)n);
    size_t o_st = o_m + al(bi), o_end = o_st + al(bi);
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, o_end, (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_xyz, xyz, b3)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_pt, normals_phi_theta, b2)) return rc;
    NormalsArgs A{};
    A.xyz = (const double*)(d + o_xyz); A.n = n; A.r = pixels_ray; A.eps_lmmin = 1e-10; A.penalty_mode = penalty_mode;
    A.mode = 1; A.eval_level = level; A.phi_theta = (const double*)(d + o_pt);
    A.cost = (double*)(d + o_c); A.m_out = (int32_t*)(d + o_m); A.status = (int32_t*)(d + o_st);
    if (int rc = run_normals(ctx, A)) return rc;
    if (int rc = fm3d_d2h(ctx, cost, d + o_c, sizeof(double) * (size_t)n)) return rc;
    if (int rc = fm3d_d2h(ctx, m, d + o_m, bi)) return rc;
    if (int rc = fm3d_d2h(ctx, status, d + o_st, bi)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_sweep_normals_dev(fm3d_ctx* ctx, const double* xyz, const double* center_phi_theta, int n,
                           int pixels_ray, int level, int penalty_mode, int n_phi, int n_theta,
                           double dphi, double dtheta, double* cost, int32_t* best_idx,
                           double* best_cost, int32_t* status) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (xyz && status)));
    FM3D_CHECK_ARG(ctx, penalty_mode >= 0 && penalty_mode <= 2);
    FM3D_CHECK_ARG(ctx, n_phi >= 1 && n_theta >= 1 && (long long)n_phi * n_theta <= (1 << 20));
    if (!ctx->has_cam || !ctx->has_g12) return fm3d_fail(ctx, FM3D_ERR_STATE, "camera / g12 not set");
    if (!ctx->has_images) return fm3d_fail(ctx, FM3D_ERR_STATE, "images not set");
    FM3D_CHECK_ARG(ctx, level >= 0 && level <= ctx->pyr.levels);
    if (pixels_ray < 0 || pixels_ray > MAX_RAY) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "pixels_ray must be in [0,%d]", MAX_RAY);
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    NormalsArgs A{};
    A.xyz = xyz; A.n = n; A.r = pixels_ray; A.eps_lmmin = 1e-10; A.penalty_mode = penalty_mode;
    A.mode = 2; A.eval_level = level; A.phi_theta = center_phi_theta;
    A.sweep_nphi = n_phi; A.sweep_ntheta = n_theta; A.sweep_dphi = dphi; A.sweep_dtheta = dtheta;
    A.cost = cost; A.best_idx = best_idx; A.best_cost = best_cost; A.status = status;
    return run_normals_fast(ctx, A);
}

int fm3d_sweep_normals(fm3d_ctx* ctx, const double* xyz, const double* center_phi_theta, int n,
                       int pixels_ray, int level, int penalty_mode, int n_phi, int n_theta,
                       double dphi, double dtheta, double* cost, int32_t* best_idx,
                       double* best_cost, int32_t* status) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && (n == 0 || (xyz && status)));
    FM3D_CHECK_ARG(ctx, n_phi >= 1 && n_theta >= 1 && (long long)n_phi * n_theta <= (1 << 20));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t K = (size_t)n_phi * n_theta;
    const size_t b3 = sizeof(double) * 3 * (size_t)n, b2 = sizeof(double) * 2 * (size_t)n, bi = sizeof(int32_t) * (size_t)n;
    const size_t bc = cost ? sizeof(double) * K * (size_t)n : 0, bd = sizeof(double) * (size_t)n;
    size_t o_xyz = 0, o_pt = o_xyz + al(b3), o_c = o_pt + al(b2), o_bi = o_c + al(bc), o_bc = o_bi + al(bi);
    size_t o_st = o_bc + al(bd), o_end = o_st + al(bi);
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, o_end, (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_xyz, xyz, b3)) return rc;
    if (center_phi_theta) if (int rc = fm3d_h2d(ctx, d + o_pt, center_phi_theta, b2)) return rc;
    if (int rc = fm3d_sweep_normals_dev(ctx, (const double*)(d + o_xyz), center_phi_theta ? (const double*)(d + o_pt) : nullptr, n,
                                        pixels_ray, level, penalty_mode, n_phi, n_theta, dphi, dtheta,
                                        cost ? (double*)(d + o_c) : nullptr, (int32_t*)(d + o_bi), (double*)(d + o_bc),
                                        (int32_t*)(d + o_st))) return rc;
    if (cost) if (int rc = fm3d_d2h(ctx, cost, d + o_c, bc)) return rc;
    if (best_idx) if (int rc = fm3d_d2h(ctx, best_idx, d + o_bi, bi)) return rc;
    if (best_cost) if (int rc = fm3d_d2h(ctx, best_cost, d + o_bc, bd)) return rc;
    if (int rc = fm3d_d2h(ctx, status, d + o_st, bi)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

int fm3d_feature_frames_dev(fm3d_ctx* ctx, const double* xyz, const double* normals, int n,
                            const double gravity[3], double* frames) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && gravity && (n == 0 || (xyz && normals && frames)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    frames_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(xyz, normals, n, gravity[0], gravity[1], gravity[2], frames);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_feature_frames(fm3d_ctx* ctx, const double* xyz, const double* normals, int n,
                        const double gravity[3], double* frames) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && gravity && (n == 0 || (xyz && normals && frames)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t b3 = sizeof(double) * 3 * (size_t)n, bf = sizeof(double) * 16 * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, 2 * al(b3) + al(bf), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, xyz, b3)) return rc;
    if (int rc = fm3d_h2d(ctx, d + al(b3), normals, b3)) return rc;
    if (int rc = fm3d_feature_frames_dev(ctx, (const double*)d, (const double*)(d + al(b3)), n, gravity,
                                         (double*)(d + 2 * al(b3)))) return rc;
    if (int rc = fm3d_d2h(ctx, frames, d + 2 * al(b3), bf)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

}  // extern "C"
