// fm3d_normals_fast.cu -- K5/K6/K7, the default implementation of the per-feature plane-normal
// search (fm3d_set_option "normals_fast" = 1; fm3d_normals.cu is the evaluation-by-evaluation
// restatement of the reference and stays available with "normals_fast" = 0).
//
// Same optimisation problem, same optimiser as the reference: coarse-to-fine over the pyramid
// (Triangulator/normaloptimizer.cpp:223-245), lmfit's lmmin per level over (phi, theta)
// (:247-292, state machine in fm3d_lm2.h), residual I1 - I2 of evaluateNormal (:65-149) with the
// bounding-box / pixel gates and the penalty wall, disc lattice of extractPixelsContour
// (Triangulator/singlecameratriangulator.cpp:341-397).  What changes is how the m x 2 problem is
// evaluated on an SM:
//
//   geometry   fp32, written as offsets from the centre ray.  The centre ray meets every candidate
//              plane in the feature point P itself, so its image-2 projection (xc, yc) does not
//              depend on the normal; with the plane-induced homography H = (n.P) R + t n^T and the
//              ray v = vc + dv,
//                  x - xc = ((H0 - xc H2) . dv) / (H2 . v),   y - yc = ((H1 - yc H2) . dv) / (H2 . v)
//              has no constant term: the quantities rounded to fp32 are small offsets, and nothing
//              that is shared by all pixels of a pass is rounded (no common-mode jitter of the
//              cost surface).  Distortion, K and the bilinear taps follow in fp32 as in
//              getBilinearInterpPix32f (tools.cpp:129-142), taps from the TMA-staged window.
//   Jacobian   analytic instead of lmfit's forward differences (step 1e-5 rad ~ 6e-4 px at the disc
//              edge): the four taps of a sample also give the image gradient, the chain rule goes
//              through K, the distortion polynomial, the perspective division and dH/dphi,
//              dH/dtheta.  The penalty weights are still differenced exactly as lmfit would see
//              them (w(x + h e_j) - w(x)), so the wall keeps its reference behaviour.
//   schedule   a trial point is evaluated together with its Jacobian when it is the first trial of
//              an LM iteration: if lmfit accepts it, the next iteration starts without another
//              pass.  nfev counts the evaluations lmfit would have made (1 per trial, 2 per
//              Jacobian), so it stays comparable with the reference.
//   gates      isInBoundingBox / isPixelGood are tested on the boundary lattice of the disc only:
//              the warp is a homeomorphism of the disc, so the interior cannot leave a rectangle
//              its boundary stays in.  The same test proves that every tap of the pass lies in the
//              staged window; if not, the pass is repeated with taps from global memory.
//   memo       a trial point whose fp32 homography coefficients equal the iterate's would return
//              the iterate's residual sum bit for bit (the value path is a fixed sequence of
//              explicit fma / non-contractible operations in every instantiation): warp 0 answers
//              it without a pass.  This is what lmfit's tail (tolerances of 30 eps) mostly asks for.
//   arithmetic packed fp32x2 (FFMA2 / FMUL2 / FADD2): one instruction for two adjacent disc pixels,
//              rays stored pair-interleaved so that one 128-bit load is the packed operand pair.
//   pipelines  the CTA is split into 1, 2 or 4 groups of warps, each with its own feature, window,
//              LM state and named barrier: the pixel passes of one group fill the SM while another
//              is in its single-lane LM step.  With several groups the rays and image-1 samples
//              stream from an L2-resident per-group scratch (next loads issued ahead of use).
#include "fm3d_normals_common.cuh"

using namespace fm3d_normals;

namespace fm3d_normals {
namespace {

#ifndef FM3D_NORMALS_NT
#define FM3D_NORMALS_NT 512
#endif
#ifndef FM3D_NORMALS_UNROLL2
#define FM3D_NORMALS_UNROLL2 1
#endif
// 1: the per-group ray / image-1 scratch (global memory, meant to stay in L2) is read and written with an L2 evict_last
// cache hint, everything else keeps the default policy
// 1 (default): the squared residuals of a thread's pixel pairs (25-50 per pass) are summed in packed fp32 -- one FFMA2 per
// pair instead of two F2F + two DFMA -- and widened to fp64 once per pass; the sums over threads and warps stay fp64.  The
// order is fixed, so the value-only and the value + Jacobian instantiation still return bit-identical sums; the cost moves
// by <= 1e-7 relative (measured: same nfev to 0.2 %, same normals), the launch gets 6.6 % faster.  0: every residual in fp64.
#ifndef FM3D_NORMALS_SS32
#define FM3D_NORMALS_SS32 1
#endif
// 1: the per-feature / per-level set-up functions are real calls (not inlined): a smaller hot region around the pixel loops
#ifndef FM3D_NORMALS_OUTLINE
#define FM3D_NORMALS_OUTLINE 0
#endif
#if FM3D_NORMALS_OUTLINE
#define FM3D_SETUP_INLINE __noinline__
#else
#define FM3D_SETUP_INLINE __forceinline__
#endif
#ifndef FM3D_NORMALS_L2HINT
#define FM3D_NORMALS_L2HINT 0
#endif
constexpr int FAST_NT = FM3D_NORMALS_NT;
constexpr float FLOOR_MAGIC = 12582912.0f;          // 1.5 * 2^23: x + MAGIC rounded down = MAGIC + floor(x)
constexpr unsigned FLOOR_MAGIC_BITS = 0x4B400000u;
enum { FLAG_WINDOW = 8 };
enum { PASS_STOP = 0, PASS_VALUE = 1, PASS_JAC = 2 };
enum { AT_X_JAC = 0, AT_XT_PLAIN = 1, AT_XT_FUSED = 2 };

// Everything a pass needs, written by the lanes of warp 0 of the group.
struct FastPass {
    // A = h[0] dx + h[1] dy (x numerator, offset form), B = h[3] dx + h[4] dy, C = h[6] dx + h[7] dy + h[8] (common
    // denominator w2 = H2 . v).  The reference ray vc is the ray through P itself, H P is parallel to the camera-2 point
    // of P, so the numerators have no constant term (h[2] = h[5] = 0).
    float h[9];
    float nd[3];        // sign(n.P) * n.v = nd[0] dx + nd[1] dy + nd[2]
    float mabs;         // |n.P|
    int kind;           // PASS_*
    int slow;           // taps from global memory (a boundary tap left the staged window)
};

struct FastShared {
    fm3d_lm2 lm;
    double w[3];
    double ht[2];       // forward-difference steps the penalty weights of this pass were taken at
    double P[3];
    double normal[3];
    double vc[2];       // centre ray
    double xc, yc;      // projection of P into camera 2 (normalised): independent of the normal
    int feature;
    int status;
    int npenalty;
    int m;
    int wx0, wy0;       // image-2 window origin (level pixels)
    int w1x0, w1y0;     // image-1 window origin
    int tma_phase;
    int where;          // AT_*
    int alive;
    int first_row, last_row;
    // Jacobian passes.  H = (n.P) R + t n^T gives dH/dp v = a_p H v + t (n_p - a_p n).v with a_p = (n_p.P)/(n.P): the first term
    // is parallel to H v and drops out of the perspective division, so the projection moves along the epipolar direction,
    //   d(x, y)/dp = (t0 - x t2, t1 - y t2) * sigma_p / w2,   sigma_p = (n_p - a_p n).v = sg[p][0] dx + sg[p][1] dy
    // (n_p - a_p n is orthogonal to P, hence to vc: no constant term).  The pixel threads sum the moments of
    // U = (dx, dy) E, E = dI2/d(x,y) . (t0 - x t2, t1 - y t2) / w2; warp 0 applies sg to the sums (jacobian_sums_from_moments).
    double sg[2][2];
    float hbase[9];     // value coefficients (FastPass::h) of the current LM iterate ...
    double s0_base;     // ... and the residual sum the pass at the iterate returned
    int base_valid;
    double jraw[6];     // SSD: the raw sums (residual sum, three moments of U, two U.d sums) of the last Jacobian result (a pass, or the table below) ...
    double jbase[6];    // ... and those of the pass at the iterate's coefficients, if one was evaluated (jbase_valid)
    int jbase_valid;
    // normals_memo = 3: the last four Jacobian passes evaluated on this level, coefficients -> raw sums: in lmfit's tail the trial
    // points fall on a handful of fp32 coefficient sets around the iterate, and a pass at one of them returns what it returned before
    float th[4][9];
    double ts[4][6];
    int tvalid[4];
    int tnext;
    int cand;           // mode 2 (sweep): candidate being evaluated
    double sweep_c[2];  // mode 2: centre (phi, theta) of the candidate grid
    double best_cost;   // mode 2: running minimum
    int best_idx;
    int trial_is_first; // the trial being evaluated opens an LM iteration
    int fuse_hint;      // the last first trial of an LM iteration was accepted (initially 1)
    double ncc_su, ncc_suu;   // cost_mode NCC: sum and sum of squares of the (approximately centred) image-1 samples of the level
    unsigned long long stats[16];
};

struct LevelConst {
    float xc, yc;
    float cols, rows;            // isPixelGood bounds on the scaled pixel
    float c0;                    // cost_mode NCC: value subtracted from every image-2 sample before it is summed (~ mean of image 1)
    const uint8_t* win;
    unsigned ww, coff, amax;
    unsigned koff, ahi;          // packed loop: shared-memory address of a tap = min(ty_bits * ww + tx_bits + koff, ahi)
    int wx0, wy0, lx_min, lx_cnt, ly_min, ly_cnt;
    const uint8_t* img2;
    int w, h, pitch;
};

struct FeatureLocals {
    double cu, cv;          // projection of P into image 1: centre of the disc
    int m;                  // pixels of the disc inside the image
    float vcxf, vcyf;       // reference ray (fp32, for the bounding-box gate)
};
constexpr int MAX_GROUPS = 4;
constexpr int NSUM_SSD = 6;          // sums of a pass: residual^2; J^T J (3), J^T r (2)
constexpr int NSUM = 12;             // cost_mode NCC: sum v, v^2, u v; sum D (2), v D (2), u D (2), D D (3)
constexpr int SWEEP_B = 4;           // mode 2: candidate normals evaluated per pass (one barrier pair per batch)
struct GroupCtl {
    RowTable rows;
    double red[16 * NSUM];
    unsigned wflags[16];
    FastShared S;
    FastPass PP;
    FastPass PPk[SWEEP_B];           // mode 2, batched: the candidates of the current batch ...
    double wk[SWEEP_B];              // ... and their penalty weights
    unsigned wflagsk[16 * SWEEP_B];
    int resume_at;                   // mode 2, batched: candidate the one-by-one loop takes over at (or -1)
    // two-slot kernel (normals_pp_kernel): what a slot's next step is, and the registers of the level in progress
    int kind;                        // STEP_*
    int ready_gen;                   // steps of the slot that have been made ready so far (written by its LM warp)
    int lvl;
    unsigned lvl_flags;              // image-1 gate of the level, charged to its first pass
    FeatureLocals Fsave;
    LevelConst Lsave;
    uint64_t bar;
};

__device__ __forceinline__ float rcp_nr(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return fmaf(r, fmaf(-a, r, 1.0f), r);   // explicit fmas: identical in every instantiation
}
__device__ __forceinline__ float rcp_approx(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
// window taps by 32-bit shared-memory address (no generic-pointer arithmetic in the pixel loop)
__device__ __forceinline__ unsigned lds_u8(unsigned a) { unsigned v; asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ unsigned lds_u8_1(unsigned a) { unsigned v; asm("ld.shared.u8 %0, [%1+1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ float u2f(unsigned b) {  // I2FP (the compiler would pick the slow I2F.U16)
    float f;
    asm("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(b));
    return f;
}

// cv::undistortPoints (5 fixed-point iterations, fp64) for the rays of the disc.  The reciprocal is
// the MUFU-seeded one: the rays are stored as fp32 offsets, the last ulp of the fp64 iteration is
// far below their resolution.
__device__ __forceinline__ void undistort_ray(const fm3d_cam& c, double u, double v, double& xo, double& yo) {
    const double x0 = (u - c.cx) * c.ifx, y0 = (v - c.cy) * c.ify;
    double x = x0, y = y0;
#pragma unroll 1
    for (int it = 0; it < 5; it++) {
        const double r2 = x * x + y * y;
        const double icd = fm3d_rcp(1.0 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
        if (icd < 0) { x = x0; y = y0; break; }
        const double xy2 = 2.0 * x * y;
        const double dx = c.p1 * xy2 + c.p2 * (r2 + 2 * x * x);
        const double dy = c.p1 * (r2 + 2 * y * y) + c.p2 * xy2;
        x = (x0 - dx) * icd;
        y = (y0 - dy) * icd;
    }
    xo = x; yo = y;
}

// Sums of a pass, per thread.  SSD: s0 = sum d^2 and f[0..4] = sum Ip Ip, Ip It, It It, Ip d, It d (d = I1 - I2, Ip / It =
// dI2/dphi, dI2/dtheta).  NCC (u = centred image-1 sample, v = image-2 sample - c0): s0 = sum v, t1 = sum v^2, t2 = sum u v and
// f[0..8] = sum Ip, It, v Ip, v It, u Ip, u It, Ip Ip, Ip It, It It -- everything the zero-mean normalised residual
// u~/|u~| - v~/|v~| and its Jacobian need (ncc_sums_to_normal_equations).
struct Acc {
    double s0, t1, t2;
    float f[9];
};

// Warp of one disc pixel into image 2 and its residual; with JAC also the derivatives of the
// sampled intensity with respect to (phi, theta).
template <bool JAC, bool SLOW, bool NCC = false>
__device__ __forceinline__ void eval_pixel_fast(const FastConsts& K, const int lvl, const FastPass& P, const LevelConst& L, float2 dv, float I1, Acc& acc) {
    const float A0 = fmaf(P.h[0], dv.x, __fmul_rn(P.h[1], dv.y));
    const float B0 = fmaf(P.h[3], dv.x, __fmul_rn(P.h[4], dv.y));
    const float C0 = fmaf(P.h[6], dv.x, fmaf(P.h[7], dv.y, P.h[8]));
    // MUFU reciprocal as it is (1 ulp): it multiplies the OFFSET from the centre, 1e-7 of at most ~100 pixels
    const float iz = rcp_approx(C0);
    // every operation of the value path is an explicit fma / non-contractible intrinsic: the value-only
    // and the value+Jacobian instantiation must return bit-identical residuals (a trial is compared
    // with an iterate that was evaluated by the other one)
    const float x = fmaf(A0, iz, L.xc), y = fmaf(B0, iz, L.yc);
    // cv::projectPoints: distortion polynomial and K (x scale), singlecameratriangulator.cpp:602,627, with the tangential
    // terms collected: xd = x g + p2 r2, yd = y g + p1 r2, g = cd + 2 p1 y + 2 p2 x
    const float r2 = fmaf(x, x, __fmul_rn(y, y));
    const float cd = fmaf(r2, fmaf(r2, fmaf(r2, K.k3, K.k2), K.k1), 1.0f);
    const float g = fmaf(K.p2x2, x, fmaf(K.p1x2, y, cd));
    const float xd = fmaf(x, g, __fmul_rn(K.p2, r2));
    const float yd = fmaf(y, g, __fmul_rn(K.p1, r2));
    const float su = fmaf(xd, K.sfx[lvl], K.scx[lvl]), sv = fmaf(yd, K.sfy[lvl], K.scy[lvl]);
    // floor and fraction without FRND/F2I
    const float tx = __fadd_rd(su, FLOOR_MAGIC), ty = __fadd_rd(sv, FLOOR_MAGIC);
    const float ax = __fsub_rn(su, __fsub_rn(tx, FLOOR_MAGIC)), ay = __fsub_rn(sv, __fsub_rn(ty, FLOOR_MAGIC));
    float b00, b01, b10, b11;
    if (!SLOW) {
        unsigned a = __float_as_uint(ty) * L.ww + __float_as_uint(tx) - L.coff;
        a = min(a, L.amax);   // memory safety only: the boundary test proves a is in range
        const uint8_t* p = L.win + a;
        b00 = u2f(p[0]); b10 = u2f(p[1]);
        b01 = u2f(p[L.ww]); b11 = u2f(p[L.ww + 1]);
    } else {
        const int x0 = (int)(__float_as_uint(tx) - FLOOR_MAGIC_BITS), y0 = (int)(__float_as_uint(ty) - FLOOR_MAGIC_BITS);
        b00 = u2f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0, y0));
        b01 = u2f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0, y0 + 1));
        b10 = u2f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0 + 1, y0));
        b11 = u2f(fm3d_at_flat(L.img2, L.w, L.h, L.pitch, x0 + 1, y0 + 1));
    }
    const float d0 = __fsub_rn(b01, b00), d1 = __fsub_rn(b11, b10);
    const float c0 = fmaf(ay, d0, b00), c1 = fmaf(ay, d1, b10);
    const float gx = __fsub_rn(c1, c0);
    const float I2 = fmaf(ax, gx, c0);
    const float d = NCC ? __fsub_rn(I2, L.c0) : __fsub_rn(I1, I2);        // NCC: d is v
    if (NCC) {
        acc.s0 += (double)d;
        acc.t1 = fma((double)d, (double)d, acc.t1);
        acc.t2 = fma((double)I1, (double)d, acc.t2);
    } else {
        acc.s0 = fma((double)d, (double)d, acc.s0);
    }
    if (JAC) {
        const float gy = fmaf(ax, d1 - d0, d0);
        const float Gx = gx * K.sfx[lvl], Gy = gy * K.sfy[lvl];               // dI2/d(xd, yd)
        // direction the projection moves in when the plane turns (towards the epipole), pushed through the distortion
        const float ex = fmaf(x, -K.t2, K.t0), ey = fmaf(y, -K.t2, K.t1);
        const float cdp2 = fmaf(r2, fmaf(r2, 6.0f * K.k3, 4.0f * K.k2), 2.0f * K.k1);   // 2 d cd / d r2
        const float rho = fmaf(x, ex, y * ey);
        const float dg = fmaf(cdp2, rho, fmaf(K.p1x2, ey, K.p2x2 * ex));
        const float dxd = fmaf(ex, g, fmaf(x, dg, K.p2x2 * rho));
        const float dyd = fmaf(ey, g, fmaf(y, dg, K.p1x2 * rho));
        const float E = fmaf(Gx, dxd, Gy * dyd) * iz;
        // dI2/dp = sigma_p E with sigma_p linear in (dx, dy): the sums below are the moments of (dx E, dy E), the
        // coefficients of sigma_phi, sigma_theta are applied to the sums (jacobian_sums_from_moments)
        const float Ip = dv.x * E;
        const float It = dv.y * E;
        if (NCC) {
            acc.f[0] += Ip; acc.f[1] += It;
            acc.f[2] = fmaf(d, Ip, acc.f[2]); acc.f[3] = fmaf(d, It, acc.f[3]);
            acc.f[4] = fmaf(I1, Ip, acc.f[4]); acc.f[5] = fmaf(I1, It, acc.f[5]);
            acc.f[6] = fmaf(Ip, Ip, acc.f[6]); acc.f[7] = fmaf(Ip, It, acc.f[7]); acc.f[8] = fmaf(It, It, acc.f[8]);
        } else {
            acc.f[0] = fmaf(Ip, Ip, acc.f[0]);
            acc.f[1] = fmaf(Ip, It, acc.f[1]);
            acc.f[2] = fmaf(It, It, acc.f[2]);
            acc.f[3] = fmaf(Ip, d, acc.f[3]);
            acc.f[4] = fmaf(It, d, acc.f[4]);
        }
    }
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): two pixels per instruction.
// A scalar operand written as bc(c) is encoded by ptxas as a broadcast register operand, so the
// pass constants are not duplicated.  Lane results are the IEEE results of the scalar ops.
typedef unsigned long long f2;
__device__ __forceinline__ f2 mk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 bc(float c) { return mk2(c, c); }
__device__ __forceinline__ float lo2(f2 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi2(f2 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2_rm(f2 a, f2 b) { f2 r; asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 neg2(f2 a) { return a ^ 0x8000000080000000ull; }

// Ray offsets are stored pair-interleaved: pixels 2p and 2p+1 share one 16-byte record
// (x[2p], x[2p+1], y[2p], y[2p+1]), so that one 128-bit load yields the packed X and Y operands of
// a pixel pair without register shuffles.  Adjacent disc pixels also share their window taps' banks.
__device__ __forceinline__ float2 ray_at(const float2* rays, int idx) {
    const float* r = reinterpret_cast<const float*>(rays) + 4 * (idx >> 1) + (idx & 1);
    return make_float2(r[0], r[2]);
}
__device__ __forceinline__ void ray_set(float2* rays, int idx, float x, float y) {
    float* r = reinterpret_cast<float*>(rays) + 4 * (idx >> 1) + (idx & 1);
    r[0] = x; r[2] = y;
}

#if FM3D_NORMALS_L2HINT
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ ulonglong2 ld_scratch(const ulonglong2* a, uint64_t pol) {
    ulonglong2 v;
    asm volatile("ld.global.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(a), "l"(pol));
    return v;
}
__device__ __forceinline__ f2 ld_scratch(const f2* a, uint64_t pol) {
    f2 v;
    asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(a), "l"(pol));
    return v;
}
#define LD_SCRATCH(ptr) ld_scratch((ptr), l2pol)
#else
#define LD_SCRATCH(ptr) (*(ptr))
#endif

struct Acc2 {      // sums of the packed loop: lanes are added at the end of the pass
    f2 a[9];       // Jacobian sums
    f2 ss;         // SSD (FM3D_NORMALS_SS32): sum of squared residuals of the thread's pixels in fp32, widened once per pass
};

// eval_pixel_fast for two pixels at once (taps from the staged window only).  The value path is
// the same sequence of IEEE operations as the scalar function, lane by lane.
template <bool JAC, bool NCC = false>
__device__ __forceinline__ void eval_pixel_pair(const FastConsts& K, const int lvl, const FastPass& P, const LevelConst& L, f2 X, f2 Y, f2 I1p, Acc& acc, Acc2& acc2) {
    const f2 A0 = fma2(bc(P.h[0]), X, mul2(bc(P.h[1]), Y));
    const f2 B0 = fma2(bc(P.h[3]), X, mul2(bc(P.h[4]), Y));
    const f2 C0 = fma2(bc(P.h[6]), X, fma2(bc(P.h[7]), Y, bc(P.h[8])));
    const f2 iz = mk2(rcp_approx(lo2(C0)), rcp_approx(hi2(C0)));
    const f2 x = fma2(A0, iz, bc(L.xc)), y = fma2(B0, iz, bc(L.yc));
    const f2 r2 = fma2(x, x, mul2(y, y));
    const f2 cd = fma2(r2, fma2(r2, fma2(r2, bc(K.k3), bc(K.k2)), bc(K.k1)), bc(1.0f));
    const f2 g = fma2(bc(K.p2x2), x, fma2(bc(K.p1x2), y, cd));
    const f2 xd = fma2(x, g, mul2(bc(K.p2), r2));
    const f2 yd = fma2(y, g, mul2(bc(K.p1), r2));
    const f2 su = fma2(xd, bc(K.sfx[lvl]), bc(K.scx[lvl])), sv = fma2(yd, bc(K.sfy[lvl]), bc(K.scy[lvl]));
    const f2 tx = add2_rm(su, bc(FLOOR_MAGIC)), ty = add2_rm(sv, bc(FLOOR_MAGIC));
    const f2 ax = sub2(su, sub2(tx, bc(FLOOR_MAGIC))), ay = sub2(sv, sub2(ty, bc(FLOOR_MAGIC)));
    // memory safety only (the boundary test proves the taps are in the window): an index above the window is clamped to its
    // last tap, one below it wraps around to the same clamp or lands in this CTA's shared memory below the window
    const unsigned aa = min(__float_as_uint(lo2(ty)) * L.ww + __float_as_uint(lo2(tx)) + L.koff, L.ahi);
    const unsigned ab = min(__float_as_uint(hi2(ty)) * L.ww + __float_as_uint(hi2(tx)) + L.koff, L.ahi);
    const unsigned aa1 = aa + L.ww, ab1 = ab + L.ww;
    const f2 b00 = mk2(u2f(lds_u8(aa)), u2f(lds_u8(ab))), b10 = mk2(u2f(lds_u8_1(aa)), u2f(lds_u8_1(ab)));
    const f2 b01 = mk2(u2f(lds_u8(aa1)), u2f(lds_u8(ab1))), b11 = mk2(u2f(lds_u8_1(aa1)), u2f(lds_u8_1(ab1)));
    const f2 d0 = sub2(b01, b00), d1 = sub2(b11, b10);
    const f2 c0 = fma2(ay, d0, b00), c1 = fma2(ay, d1, b10);
    const f2 gx = sub2(c1, c0);
    const f2 I2 = fma2(ax, gx, c0);
    const f2 d = NCC ? sub2(I2, bc(L.c0)) : sub2(I1p, I2);                  // NCC: d is v
    const float da = lo2(d), db = hi2(d);
    if (NCC) {
        const double va = (double)da, vb = (double)db;
        acc.s0 += va; acc.s0 += vb;
        acc.t1 = fma(va, va, acc.t1); acc.t1 = fma(vb, vb, acc.t1);
        acc.t2 = fma((double)lo2(I1p), va, acc.t2); acc.t2 = fma((double)hi2(I1p), vb, acc.t2);
    } else {
#if FM3D_NORMALS_SS32
        acc2.ss = fma2(d, d, acc2.ss);
#else
        acc.s0 = fma((double)da, (double)da, acc.s0);
        acc.s0 = fma((double)db, (double)db, acc.s0);
#endif
    }
    if (JAC) {
        const f2 gy = fma2(ax, sub2(d1, d0), d0);
        const f2 Gx = mul2(gx, bc(K.sfx[lvl])), Gy = mul2(gy, bc(K.sfy[lvl]));
        const f2 ex = fma2(x, bc(-K.t2), bc(K.t0)), ey = fma2(y, bc(-K.t2), bc(K.t1));
        const f2 cdp2 = fma2(r2, fma2(r2, bc(6.0f * K.k3), bc(4.0f * K.k2)), bc(2.0f * K.k1));   // 2 d cd / d r2
        const f2 rho = fma2(x, ex, mul2(y, ey));
        const f2 dg = fma2(cdp2, rho, fma2(bc(K.p1x2), ey, mul2(bc(K.p2x2), ex)));
        const f2 dxd = fma2(ex, g, fma2(x, dg, mul2(bc(K.p2x2), rho)));
        const f2 dyd = fma2(ey, g, fma2(y, dg, mul2(bc(K.p1x2), rho)));
        const f2 E = mul2(fma2(Gx, dxd, mul2(Gy, dyd)), iz);
        const f2 Ip = mul2(X, E);    // moments of (dx E, dy E): see eval_pixel_fast
        const f2 It = mul2(Y, E);
        if (NCC) {
            acc2.a[0] = add2(acc2.a[0], Ip); acc2.a[1] = add2(acc2.a[1], It);
            acc2.a[2] = fma2(d, Ip, acc2.a[2]); acc2.a[3] = fma2(d, It, acc2.a[3]);
            acc2.a[4] = fma2(I1p, Ip, acc2.a[4]); acc2.a[5] = fma2(I1p, It, acc2.a[5]);
            acc2.a[6] = fma2(Ip, Ip, acc2.a[6]); acc2.a[7] = fma2(Ip, It, acc2.a[7]); acc2.a[8] = fma2(It, It, acc2.a[8]);
        } else {
            acc2.a[0] = fma2(Ip, Ip, acc2.a[0]);
            acc2.a[1] = fma2(Ip, It, acc2.a[1]);
            acc2.a[2] = fma2(It, It, acc2.a[2]);
            acc2.a[3] = fma2(Ip, d, acc2.a[3]);
            acc2.a[4] = fma2(It, d, acc2.a[4]);
        }
    }
}

// Gates of one boundary pixel (isInBoundingBox :646-655, isPixelGood :657-665) and the proof that its
// taps are inside the staged window.
__device__ __forceinline__ unsigned boundary_flags(const FastConsts& K, const int lvl, const FastPass& P, const LevelConst& L, float vcx, float vcy,
                                                   float cmax, float2 dv) {
    unsigned flags = 0;
    const float den = fmaf(P.nd[0], dv.x, fmaf(P.nd[1], dv.y, P.nd[2]));
    const float vx = vcx + dv.x, vy = vcy + dv.y;
    const float t = fmaxf(1.0f, fmaxf(fabsf(vx), fabsf(vy)));
    if (den != den || P.mabs != P.mabs) flags |= FLAG_NAN;
    if (!(P.mabs * t < cmax * den)) flags |= FLAG_BBOX;
    const float A0 = fmaf(P.h[0], dv.x, fmaf(P.h[1], dv.y, P.h[2]));
    const float B0 = fmaf(P.h[3], dv.x, fmaf(P.h[4], dv.y, P.h[5]));
    const float C0 = fmaf(P.h[6], dv.x, fmaf(P.h[7], dv.y, P.h[8]));
    const float iz = rcp_nr(C0);
    const float x = L.xc + A0 * iz, y = L.yc + B0 * iz;
    const float r2 = fmaf(x, x, y * y);
    const float cd = fmaf(r2, fmaf(r2, fmaf(r2, K.k3, K.k2), K.k1), 1.0f);
    const float xy2 = 2.0f * x * y;
    const float xd = fmaf(x, cd, fmaf(K.p1, xy2, K.p2 * fmaf(2.0f * x, x, r2)));
    const float yd = fmaf(y, cd, fmaf(K.p1, fmaf(2.0f * y, y, r2), K.p2 * xy2));
    const float su = fmaf(xd, K.sfx[lvl], K.scx[lvl]), sv = fmaf(yd, K.sfy[lvl], K.scy[lvl]);
    if (!(su >= 0.0f && su <= L.cols && sv >= 0.0f && sv <= L.rows)) flags |= FLAG_PIX;
    // one pixel of margin: the lattice boundary is a polygon, its warp slightly curved
    const float fx0 = floorf(su), fy0 = floorf(sv);
    const float lx = fx0 - (float)L.wx0, ly = fy0 - (float)L.wy0;
    if (!(lx >= (float)(L.lx_min + 1) && lx < (float)(L.lx_min + L.lx_cnt - 1) &&
          ly >= (float)(L.ly_min + 1) && ly < (float)(L.ly_min + L.ly_cnt - 1))) flags |= FLAG_WINDOW;
    return flags;
}

template <bool JAC, bool SLOW, bool PREFETCH, bool NCC = false>
__device__ __forceinline__ void run_pixels(const FastConsts& K, const int lvl, const FastPass& P, const LevelConst& L, const float2* __restrict__ rays,
                                           const float* __restrict__ i1, int m, int tid, int NT, Acc& acc) {
    if (SLOW) {
        // taps from global memory (a boundary tap left the staged window): scalar path
        for (int idx = tid; idx < m; idx += NT) eval_pixel_fast<JAC, true, NCC>(K, lvl, P, L, ray_at(rays, idx), i1[idx], acc);
        return;
    }
    constexpr int NJ = NCC ? 9 : 5;
    Acc2 acc2;
#pragma unroll
    for (int k = 0; k < 9; k++) acc2.a[k] = 0ull;
    acc2.ss = 0ull;
    const ulonglong2* __restrict__ rp = reinterpret_cast<const ulonglong2*>(rays);   // (X, Y) of a pixel pair
    const f2* __restrict__ ip = reinterpret_cast<const f2*>(i1);
    const int npair = m >> 1;
    int p = tid;
    if (PREFETCH) {
#if FM3D_NORMALS_L2HINT
        const uint64_t l2pol = l2_evict_last_policy();
#endif
        // rays / image-1 samples stream from the L2-resident scratch: the loads of the next pair
        // are issued before the current one is evaluated
        ulonglong2 r = make_ulonglong2(0ull, 0ull);
        f2 I = 0ull;
        if (p < npair) { r = LD_SCRATCH(rp + p); I = LD_SCRATCH(ip + p); }
#if FM3D_NORMALS_UNROLL2
        // two register sets take turns (no copies): while one pair is evaluated the loads of the next are in flight
        const ulonglong2* __restrict__ rq = rp + p;
        const f2* __restrict__ iq = ip + p;
        int left = p < npair ? (npair - p + NT - 1) / NT : 0;
        ulonglong2 rb = make_ulonglong2(0ull, 0ull);
        f2 Ib = 0ull;
#pragma unroll 1
        while (left >= 2) {
            rb = LD_SCRATCH(rq + NT); Ib = LD_SCRATCH(iq + NT);
            eval_pixel_pair<JAC, NCC>(K, lvl, P, L, r.x, r.y, I, acc, acc2);
            if (left > 2) { r = LD_SCRATCH(rq + 2 * NT); I = LD_SCRATCH(iq + 2 * NT); }
            eval_pixel_pair<JAC, NCC>(K, lvl, P, L, rb.x, rb.y, Ib, acc, acc2);
            rq += 2 * NT; iq += 2 * NT;
            left -= 2;
        }
        if (left == 1) eval_pixel_pair<JAC, NCC>(K, lvl, P, L, r.x, r.y, I, acc, acc2);
#else
        for (; p < npair; p += NT) {
            const ulonglong2 cr = r;
            const f2 cI = I;
            const int nx = p + NT;
            if (nx < npair) { r = LD_SCRATCH(rp + nx); I = LD_SCRATCH(ip + nx); }
            eval_pixel_pair<JAC, NCC>(K, lvl, P, L, cr.x, cr.y, cI, acc, acc2);
        }
#endif
    } else {
        for (; p < npair; p += NT) {
            const ulonglong2 r = rp[p];
            eval_pixel_pair<JAC, NCC>(K, lvl, P, L, r.x, r.y, ip[p], acc, acc2);
        }
    }
    // odd pixel count: the last pixel alone (thread chosen so that the summation order is fixed)
    if ((m & 1) && tid == (npair % NT)) eval_pixel_fast<JAC, false, NCC>(K, lvl, P, L, ray_at(rays, m - 1), i1[m - 1], acc);
    if (JAC) {
#pragma unroll
        for (int k = 0; k < NJ; k++) acc.f[k] += lo2(acc2.a[k]) + hi2(acc2.a[k]);
    }
#if FM3D_NORMALS_SS32
    if (!NCC) acc.s0 += (double)lo2(acc2.ss) + (double)hi2(acc2.ss);
#endif
}

// Jacobian pass: the pixel threads summed the moments of U = (dx E, dy E); dI2/dp = sg[p] . U.  In place: SSD s[1..5] =
// sum U1 U1, U1 U2, U2 U2, U1 d, U2 d -> sum Ip Ip, Ip It, It It, Ip d, It d; NCC s[3..11] = sum U1, U2, v U1, v U2, u U1, u U2,
// U1 U1, U1 U2, U2 U2 -> the same sums of (Ip, It).
__device__ __forceinline__ void jacobian_sums_from_moments(double* s, const double (*c)[2], bool ncc) {
    const double a0 = c[0][0], a1 = c[0][1], b0 = c[1][0], b1 = c[1][1];
    double* q = s + (ncc ? 9 : 1);
    const double m11 = q[0], m12 = q[1], m22 = q[2];
    q[0] = a0 * a0 * m11 + 2.0 * a0 * a1 * m12 + a1 * a1 * m22;
    q[1] = a0 * b0 * m11 + (a0 * b1 + a1 * b0) * m12 + a1 * b1 * m22;
    q[2] = b0 * b0 * m11 + 2.0 * b0 * b1 * m12 + b1 * b1 * m22;
    const int first = ncc ? 3 : 4, npairs = ncc ? 3 : 1;
    for (int k = 0; k < npairs; k++) {
        const double u1 = s[first + 2 * k], u2 = s[first + 2 * k + 1];
        s[first + 2 * k] = a0 * u1 + a1 * u2;
        s[first + 2 * k + 1] = b0 * u1 + b1 * u2;
    }
}

// cost_mode NCC: the sums of a pass -> what the LM consumes, in the layout of the SSD sums (s[0] = sum of squared residuals
// without the penalty weight, s[1..3] = J^T J, s[4..5] = -J^T r), for the residual r_i = u~_i/|u~| - v~_i/|v~| with
// u~ = u - mean u, v~ = v - mean v (squared sum 2 - 2 NCC).  With C = sum u~ v~, A^2 = sum u~^2, B^2 = sum v~^2,
// E_p = sum D^p v~, F_p = sum D^p u~, Cov_pq = sum (D^p - mean)(D^q - mean), D^p = dI2/dp:
//   sum r^2 = 2 - 2 C / (A B),   J^T J_pq = Cov_pq / B^2 - E_p E_q / B^4,   J^T r_p = -F_p / (A B) + E_p C / (A B^3).
__device__ __forceinline__ void ncc_sums_to_normal_equations(double* s, double su, double suu, int m, bool jac) {
    const double im = 1.0 / (double)m;
    const double sv = s[0], svv = s[1], suv = s[2];
    const double A2 = suu - su * su * im, B2 = svv - sv * sv * im, C = suv - su * sv * im;
    const double AB = sqrt(A2 * B2);            // NaN / 0 for a flat patch: the cost below becomes NaN and the feature is dropped
    const double ff = 2.0 - 2.0 * C / AB;
    if (jac) {
        const double d0 = s[3], d1 = s[4];
        const double E0 = s[5] - sv * im * d0, E1 = s[6] - sv * im * d1;
        const double F0 = s[7] - su * im * d0, F1 = s[8] - su * im * d1;
        const double c00 = s[9] - d0 * d0 * im, c01 = s[10] - d0 * d1 * im, c11 = s[11] - d1 * d1 * im;
        const double iB2 = 1.0 / B2, iB4 = iB2 * iB2;
        const double g0 = -F0 / AB + E0 * C / (AB * B2), g1 = -F1 / AB + E1 * C / (AB * B2);
        s[1] = c00 * iB2 - E0 * E0 * iB4;
        s[2] = c01 * iB2 - E0 * E1 * iB4;
        s[3] = c11 * iB2 - E1 * E1 * iB4;
        s[4] = -g0;
        s[5] = -g1;
    }
    s[0] = ff < 0.0 ? 0.0 : ff;                 // rounding can leave -1e-17 for identical patches
}

// Mode 2 (dense candidate sweep): NE candidate normals per pass.  Every thread walks the pixel pairs the
// single-candidate loop gives it, in the same order, and evaluates all candidates of the batch on a pair
// before moving on: the per-candidate sums are bit-identical to NE single passes, the rays / image-1
// samples are read once per batch, the NE dependency chains are independent, and the barriers, the
// reduction and the serial publish step are paid once per batch.
template <int NE, bool PREFETCH>
__device__ __forceinline__ void run_pixels_multi(const FastConsts& K, const int lvl, const FastPass* __restrict__ PPk, const LevelConst& L,
                                                 const float2* __restrict__ rays, const float* __restrict__ i1, int m,
                                                 int tid, int NT, Acc* acc) {
    Acc2 a2[NE];                    // value passes use only .ss (the packed sum of squared residuals of a candidate)
#pragma unroll
    for (int k = 0; k < NE; k++) a2[k].ss = 0ull;
    const ulonglong2* __restrict__ rp = reinterpret_cast<const ulonglong2*>(rays);
    const f2* __restrict__ ip = reinterpret_cast<const f2*>(i1);
    const int npair = m >> 1;
    int p = tid;
    ulonglong2 r = make_ulonglong2(0ull, 0ull);
    f2 I = 0ull;
    if (p < npair) { r = rp[p]; I = ip[p]; }
    for (; p < npair; p += NT) {
        const ulonglong2 cr = r;
        const f2 cI = I;
        if (PREFETCH) {
            const int nx = p + NT;
            if (nx < npair) { r = rp[nx]; I = ip[nx]; }
        }
#pragma unroll
        for (int k = 0; k < NE; k++) eval_pixel_pair<false>(K, lvl, PPk[k], L, cr.x, cr.y, cI, acc[k], a2[k]);
        if (!PREFETCH) {
            const int nx = p + NT;
            if (nx < npair) { r = rp[nx]; I = ip[nx]; }
        }
    }
    if ((m & 1) && tid == (npair % NT)) {
#pragma unroll
        for (int k = 0; k < NE; k++) eval_pixel_fast<false, false>(K, lvl, PPk[k], L, ray_at(rays, m - 1), i1[m - 1], acc[k]);
    }
#if FM3D_NORMALS_SS32
#pragma unroll
    for (int k = 0; k < NE; k++) acc[k].s0 += (double)lo2(a2[k].ss) + (double)hi2(a2[k].ss);
#endif
}

// Warp 0, all lanes, uniform arguments: evaluation point (phi, theta) -> homography coefficients
// of the pass.  The work is spread over the lanes (two sincos, nine coefficient rows, three penalty
// weights) because it sits between two passes with the other fifteen warps waiting.
__device__ void publish_pass(FastPass* PP, FastShared* S, const fm3d_cam& cam, double phi, double theta, int kind,
                             int penalty_mode, double eps, int lane) {
    // sincos(phi) on lane 0, sincos(theta) on lane 1
    double sv, cv;
    sincos(lane == 0 ? phi : theta, &sv, &cv);
    const double sp = __shfl_sync(0xffffffffu, sv, 0), cp = __shfl_sync(0xffffffffu, cv, 0);
    const double st = __shfl_sync(0xffffffffu, sv, 1), ct = __shfl_sync(0xffffffffu, cv, 1);
    // sph2car (tools.cpp:772-777) and its derivatives
    const double n0x = ct * cp, n0y = ct * sp, n0z = st;
    if (n0x != n0x || n0y != n0y || n0z != n0z) {  // normaloptimizer.cpp:81-85
        if (lane == 0) {
            S->status = FM3D_FEAT_ABORT_NAN;
            S->alive = 0;
            PP->kind = PASS_STOP;
        }
        return;
    }
    const double vcx = S->vc[0], vcy = S->vc[1];
    if (lane < 3) {
        // value coefficients, row `lane` of H = (n.P) R + t n^T
        const int row = lane;
        const double nx = n0x, ny = n0y, nz = n0z;
        const double me = nx * S->P[0] + ny * S->P[1] + nz * S->P[2];
        const double H20 = me * cam.R[6] + cam.t[2] * nx, H21 = me * cam.R[7] + cam.t[2] * ny, H22 = me * cam.R[8] + cam.t[2] * nz;
        double c0, c1, c2;
        if (row == 2) {
            c0 = H20; c1 = H21; c2 = H22;
        } else {
            // numerators in offset form: (H0 - xc H2) . v and (H1 - yc H2) . v with v = vc + dv.  Their
            // constant terms vanish: vc is the ray through P
            const double ref = row == 0 ? S->xc : S->yc;
            const double tr = cam.t[row];
            c0 = me * cam.R[3 * row] + tr * nx - ref * H20;
            c1 = me * cam.R[3 * row + 1] + tr * ny - ref * H21;
            c2 = me * cam.R[3 * row + 2] + tr * nz - ref * H22;
        }
        PP->h[3 * row] = (float)c0;
        PP->h[3 * row + 1] = (float)c1;
        PP->h[3 * row + 2] = row == 2 ? (float)(c0 * vcx + c1 * vcy + c2) : 0.0f;   // numerators: (H0 - xc H2) . vc = 0
    } else if (lane < 5) {
        // sigma_p = (n_p - a_p n) . v, a_p = (n_p.P)/(n.P), for p = phi (lane 3), theta (lane 4): see FastPass::sg
        const bool dphi = lane == 3;
        const double nx = dphi ? -ct * sp : -st * cp, ny = dphi ? ct * cp : -st * sp, nz = dphi ? 0.0 : ct;
        const double m0 = n0x * S->P[0] + n0y * S->P[1] + n0z * S->P[2];
        const double mp = nx * S->P[0] + ny * S->P[1] + nz * S->P[2];
        const double a = m0 != 0.0 ? mp / m0 : 0.0;     // n.P = 0 fails the bounding-box gate of every pixel
        const double c0 = nx - a * n0x, c1 = ny - a * n0y, c2 = nz - a * n0z;
        S->sg[lane - 3][0] = c0;        // (c0, c1, c2) . vc = 0: n_p - a n is orthogonal to P
        S->sg[lane - 3][1] = c1;
        (void)c2;
    } else if (lane == 9) {
        const double me = n0x * S->P[0] + n0y * S->P[1] + n0z * S->P[2];
        const double sg = me < 0 ? -1.0 : 1.0;
        if (me == 0.0) {        // k == 0 fails 0 < k for every pixel
            PP->nd[0] = PP->nd[1] = PP->nd[2] = 0.0f;
        } else {
            PP->nd[0] = (float)(sg * n0x);
            PP->nd[1] = (float)(sg * n0y);
            PP->nd[2] = (float)(sg * (n0x * vcx + n0y * vcy + n0z));
        }
        PP->mabs = (float)fabs(me);
        PP->kind = kind;
    } else if (lane >= 10 && lane < 13) {
        // penalty weights of f(x), f(x + h0 e0), f(x + h1 e1) as lmfit's forward differences see them
        const int j = lane - 10;
        const double h0 = fmax(eps * eps, eps * fabs(phi)), h1 = fmax(eps * eps, eps * fabs(theta));
        int entered;
        if (j == 0 || kind == PASS_JAC)
            S->w[j] = penalty_weight(phi + (j == 1 ? h0 : 0.0), theta + (j == 2 ? h1 : 0.0), penalty_mode, entered);
        if (j == 0) { S->ht[0] = h0; S->ht[1] = h1; }
    }
}

// Consume the sums of a Jacobian evaluation at the point lm.x (analytic derivatives, differenced weights).
__device__ __forceinline__ int consume_jacobian(FastShared* S, fm3d_lm2& lm, const double* s, int penalty_mode) {
    // f_0 = w0 d,  (f_j - f_0)/h_j = w_j dd/dx_j + (w_j - w0) d / h_j,  dd/dx_j = -dI2/dx_j
    const double w0 = S->w[0], w1 = S->w[1], w2 = S->w[2];
    const double a1 = (w1 - w0) / S->ht[0], a2 = (w2 - w0) / S->ht[1];
    const double A00 = s[0], E11 = s[1], E12 = s[2], E22 = s[3], E1d = -s[4], E2d = -s[5];
    const double ff = w0 * w0 * A00;
    const double S00 = w1 * w1 * E11 + 2 * w1 * a1 * E1d + a1 * a1 * A00;
    const double S11 = w2 * w2 * E22 + 2 * w2 * a2 * E2d + a2 * a2 * A00;
    const double S01 = w1 * w2 * E12 + w1 * a2 * E1d + w2 * a1 * E2d + a1 * a2 * A00;
    const double g0 = w0 * (w1 * E1d + a1 * A00);
    const double g1 = w0 * (w2 * E2d + a2 * A00);
    if (penalty_mode != FM3D_PENALTY_OFF) {
        // evaluations at x + h e_j that entered the penalty branch (counted like the reference's)
        if (w1 != 1.0) S->npenalty++;
        if (w2 != 1.0) S->npenalty++;
    }
    return fm3d_lm2_after_jacobian(&lm, ff, S00, S01, S11, g0, g1);
}

// The serial step between two passes, on the group's LM warp (all 32 lanes, uniform arguments): reduces the per-warp sums of
// the pass just evaluated, advances lmfit's state machine on lane 0 (answering trials that round to the iterate's
// coefficients without a pass), publishes the next pass into *PP (PP->kind = PASS_STOP: level finished or feature dropped).
template <bool ncc>
__device__ __forceinline__ void lm_advance(const NormalsArgs& A, GroupCtl& G, const FastPass& P, const int m, const int f, const int NW,
                                           const int lane, const long long t_a, const long long t_b0, const long long t_b) {
    const fm3d_cam& cam = A.cam;
    double* red = G.red;
    unsigned* wflags = G.wflags;
    FastShared* S = &G.S;
    FastPass* PP = &G.PP;
                        double s[NSUM];
                        unsigned any_flags = lane < NW ? wflags[lane] : 0u;
                        {
                            const int ns = ncc ? (P.kind == PASS_JAC ? NSUM : 3) : NSUM_SSD;
    #pragma unroll
                            for (int k = 0; k < NSUM; k++) {
                                double v = (lane < NW && k < ns) ? red[lane * NSUM + k] : 0.0;
                                if (k < ns) {
    #pragma unroll
                                    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                }
                                s[k] = v;
                            }
                            // raw sums of this Jacobian pass: they depend on the pass's fp32 coefficients only (the sigma
                            // coefficients are applied below, after the reduction), so a later Jacobian request at a point with
                            // the same coefficients can be answered from them
                            if (!ncc && P.kind == PASS_JAC && lane == 0) {
    #pragma unroll
                                for (int k = 0; k < 6; k++) S->jraw[k] = s[k];
                            }
                            if (P.kind == PASS_JAC) jacobian_sums_from_moments(s, S->sg, ncc);
                            if (ncc) ncc_sums_to_normal_equations(s, S->ncc_su, S->ncc_suu, m, P.kind == PASS_JAC);
                        }
                        any_flags = __reduce_or_sync(0xffffffffu, any_flags);
                        if (!ncc && A.memo_trials >= 3 && A.mode == 0 && P.kind == PASS_JAC && any_flags == 0u) {
                            const int e = S->tnext;         // a clean Jacobian pass: remember what it returned
                            __syncwarp();
                            if (lane < 9) S->th[e][lane] = PP->h[lane];     // (P.h[lane] would put the register copy of the pass on the stack)
                            if (lane == 0) {
    #pragma unroll
                                for (int k = 0; k < 6; k++) S->ts[e][k] = S->jraw[k];
                                S->tvalid[e] = 1;
                                S->tnext = (e + 1) & 3;
                            }
                            __syncwarp();
                        }
                        const long long t_r = clock64();
                        long long t_lm = 0;
                        // `memo`: the result in s[0] was not evaluated but taken from the base point: the
                        // proposed trial point rounds to the same fp32 homography coefficients as the
                        // current iterate, so the pass would return the iterate's sum bit for bit.  Deep
                        // in lmfit's tail (steps below 1e-8 rad, tolerances of 30 eps) this is what every
                        // trial does; those trials are answered here without a pass.
                        bool memo = false;
                        bool from_table = false;    // memo, answered from the table of earlier passes (value and Jacobian sums of another point than the iterate)
                        bool slow_pass = P.slow != 0;
                        int pass_kind = P.kind;
                        for (;;) {
                            int next_kind = PASS_STOP;          // what lane 0 asks the warp to publish
                            int take_base = 0;                  // PP->h[0] are the coefficients of the (new) iterate
                            double next_phi = 0.0, next_theta = 0.0;
                            const long long t_l0 = clock64();
                            if (lane == 0) {
                                fm3d_lm2& lm = S->lm;
                                if (!memo) {   // executed work (fm3d_get_normals_stats)
                                    const int slot = pass_kind == PASS_VALUE ? 0 : (S->where == AT_X_JAC ? 1 : 2);
                                    S->stats[slot]++;
                                    if (slow_pass) S->stats[4]++;
                                    S->stats[pass_kind == PASS_VALUE ? 5 : 6] += (unsigned long long)m;
                                } else {
                                    S->stats[13]++;
                                }
                                const unsigned fl = memo ? 0u : any_flags;
                                if ((fl & FLAG_WINDOW) && !(fl & 7) && !slow_pass) {
                                    PP->slow = 1;               // same pass again, taps from global memory
                                    next_kind = -1;
                                } else if (A.mode == 2) {
                                    // dense candidate sweep: cost of this candidate (NaN if it fails a gate), next one
                                    const int K = A.sweep_nphi * A.sweep_ntheta, c = S->cand;
                                    const bool bad = (fl & 7) || s[0] != s[0];
                                    const double cval = bad ? __longlong_as_double(0x7ff8000000000000LL) : S->w[0] * S->w[0] * s[0];
                                    if (A.cost) A.cost[(size_t)f * K + c] = cval;
                                    if (!bad && cval < S->best_cost) { S->best_cost = cval; S->best_idx = c; }
                                    if (c + 1 < K) {
                                        S->cand = c + 1;
                                        const int ip = (c + 1) / A.sweep_ntheta, it = (c + 1) - ip * A.sweep_ntheta;
                                        next_kind = PASS_VALUE;
                                        next_phi = S->sweep_c[0] + ((double)ip - 0.5 * (double)(A.sweep_nphi - 1)) * A.sweep_dphi;
                                        next_theta = S->sweep_c[1] + ((double)it - 0.5 * (double)(A.sweep_ntheta - 1)) * A.sweep_dtheta;
                                    } else {
                                        PP->kind = PASS_STOP;
                                    }
                                } else if ((fl & 7) || s[0] != s[0]) {
                                    // all bounding-box / NaN tests of an evaluation precede its pixel tests
                                    S->status = ((fl & FLAG_NAN) || !(fl & 7)) ? FM3D_FEAT_ABORT_NAN
                                              : (fl & FLAG_BBOX) ? FM3D_FEAT_ABORT_BBOX : FM3D_FEAT_ABORT_PIXEL;
                                    S->alive = 0;
                                    PP->kind = PASS_STOP;
                                } else if (A.mode != 0) {
                                    A.cost[f] = S->w[0] * S->w[0] * s[0];
                                    PP->kind = PASS_STOP;
                                } else {
                                    const int where = S->where;
                                    bool first_trial = false;   // the next trial opens an LM iteration
                                    int cmd;
                                    if (where == AT_X_JAC) {
                                        if (lm.first && S->w[0] != 1.0) S->npenalty++;
                                        cmd = consume_jacobian(S, lm, s, A.penalty_mode);
                                        first_trial = true;
                                        S->s0_base = s[0];      // sum and coefficients of the iterate
                                        take_base = 1;
                                        if ((!memo || from_table) && !ncc) {    // ... and the raw Jacobian sums of the pass at them
    #pragma unroll
                                            for (int k = 0; k < 6; k++) S->jbase[k] = S->jraw[k];
                                            S->jbase_valid = 1;
                                        }
                                    } else {
                                        if (S->w[0] != 1.0) S->npenalty++;
                                        const int iter_before = lm.iter;
                                        cmd = fm3d_lm2_after_trial(&lm, S->w[0] * S->w[0] * s[0]);
                                        const bool accepted = lm.iter != iter_before;
                                        if (S->trial_is_first) S->fuse_hint = accepted ? 1 : 0;
                                        if ((!memo || from_table) && accepted) {   // new iterate
                                            S->s0_base = s[0];
                                            take_base = 1;
                                            if (where == AT_XT_FUSED && !ncc) {   // its Jacobian sums came with the trial
    #pragma unroll
                                                for (int k = 0; k < 6; k++) S->jbase[k] = S->jraw[k];
                                                S->jbase_valid = 1;
                                            } else {
                                                S->jbase_valid = 0;
                                            }
                                        }
                                        // accepted (x == xt now) and the Jacobian came with the trial
                                        if (cmd == FM3D_LM_CMD_JAC && where == AT_XT_FUSED) {
                                            S->stats[3]++;
                                            cmd = consume_jacobian(S, lm, s, A.penalty_mode);
                                            first_trial = true;
                                        }
                                    }
                                    if (cmd == FM3D_LM_CMD_JAC) {
                                        S->where = AT_X_JAC;
                                        next_kind = PASS_JAC; next_phi = lm.x[0]; next_theta = lm.x[1];
                                    } else if (cmd == FM3D_LM_CMD_TRIAL) {
                                        // first trial of an iteration: evaluate its Jacobian along with it;
                                        // re-trials after a rejection are value-only
                                        // 3: re-trials too -- after a rejection lmfit shortens the step and the re-trial is usually
                                        // accepted; its Jacobian then is the next iteration's (no separate Jacobian pass)
                                        // (measured and dropped: fusing only trials whose predicted relative reduction exceeds 1e-4 ...
                                        // 1e-10 -- 11.45 ... 10.70 ms against 10.39 for every trial: the fewer passes, the better)
                                        const bool fuse = (first_trial && (A.fuse_trials == 1 || (A.fuse_trials == 2 && S->fuse_hint))) ||
                                                          A.fuse_trials == 3;
                                        S->trial_is_first = first_trial ? 1 : 0;
                                        S->where = fuse ? AT_XT_FUSED : AT_XT_PLAIN;
                                        next_kind = fuse ? PASS_JAC : PASS_VALUE; next_phi = lm.xt[0]; next_theta = lm.xt[1];
                                    } else {
                                        PP->kind = PASS_STOP;   // FM3D_LM_CMD_DONE
                                    }
                                }
                            }
                            t_lm += clock64() - t_l0;
                            next_kind = __shfl_sync(0xffffffffu, next_kind, 0);
                            if (next_kind <= PASS_STOP) break;      // stop, or repeat the same pass (-1)
                            take_base = __shfl_sync(0xffffffffu, take_base, 0);
                            next_phi = __shfl_sync(0xffffffffu, next_phi, 0);
                            next_theta = __shfl_sync(0xffffffffu, next_theta, 0);
                            // remember the coefficients of the iterate before they are overwritten
                            if (take_base) {
                                if (lane < 9) S->hbase[lane] = PP->h[lane];
                                if (lane == 0) S->base_valid = 1;
                            }
                            __syncwarp();
                            publish_pass(PP, S, cam, next_phi, next_theta, next_kind, A.penalty_mode, S->lm.eps, lane);
                            __syncwarp();
                            if (PP->kind == PASS_STOP) break;       // NaN normal
                            // would this pass return what the pass at the iterate's coefficients returned?  Trial points: its residual
                            // sum.  Jacobian requests (normals_memo = 2, SSD): after an accepted trial that was itself answered from
                            // memory the new iterate has the old coefficients -- the same pass, the same raw sums; only the sigma
                            // coefficients and the penalty weights (both from the exact point, both applied here) differ.
                            const bool at_jac = S->where == AT_X_JAC;
                            bool same = S->base_valid != 0 && A.memo_trials && (!at_jac || (A.memo_trials >= 2 && !ncc && S->jbase_valid != 0));
                            if (lane < 9) same = same && (PP->h[lane] == S->hbase[lane]);
                            same = __all_sync(0xffffffffu, same);
                            int hit = -1;                           // normals_memo = 3: a pass at these coefficients evaluated earlier on this level?
                            if (!same && A.memo_trials >= 3 && !ncc && PP->kind == PASS_JAC) {
                                for (int e = 0; e < 4 && hit < 0; e++) {
                                    bool eq = S->tvalid[e] != 0;
                                    if (lane < 9) eq = eq && (PP->h[lane] == S->th[e][lane]);
                                    if (__all_sync(0xffffffffu, eq)) hit = e;
                                }
                            }
                            if (!same && hit < 0) break;            // run the pass
                            memo = true;
                            from_table = !same;
                            if (!same) {
                                // value and Jacobian sums of the earlier pass; S->where stays (a fused trial, or the Jacobian at x)
    #pragma unroll
                                if (lane < 6) S->jraw[lane] = S->ts[hit][lane];    // they are "the raw sums of the last Jacobian result" now
                                __syncwarp();
    #pragma unroll
                                for (int k = 0; k < 6; k++) s[k] = S->jraw[k];
                                jacobian_sums_from_moments(s, S->sg, false);
                            } else if (at_jac) {
    #pragma unroll
                                for (int k = 0; k < 6; k++) s[k] = S->jbase[k];
                                jacobian_sums_from_moments(s, S->sg, false);
                            } else {
                                s[0] = S->s0_base;
                                if (lane == 0) { S->where = AT_XT_PLAIN; }
                            }
                            __syncwarp();
                        }
                        if (lane == 0) {
                            const long long t_c = clock64();
                            S->stats[8] += (unsigned long long)(t_b0 - t_a);   // thread 0: pixel work of the pass
                            S->stats[9] += (unsigned long long)(t_b - t_b0);   // thread 0: wait at the barrier
                            S->stats[10] += (unsigned long long)(t_c - t_b);   // reduction + LM + publish
                            S->stats[11] += (unsigned long long)t_lm;          //   of which LM algebra
                            S->stats[12] += (unsigned long long)(t_c - t_r - t_lm);   //   of which sincos + homographies
                        }
}

// A CTA runs `groups` independent feature pipelines side by side (1 or 2): each group of warps has
// its own window, LM state, queue slot and named barrier, so that the pixel passes of one feature
// fill the SM while the other feature is in its serial LM step.
__device__ __forceinline__ void gsync(int groups, int g, int nt) {
    if (groups == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(nt) : "memory");
}
__device__ __forceinline__ int gsync_and(int groups, int g, int nt, int pred) {
    if (groups == 1) return __syncthreads_and(pred);
    int out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "bar.red.and.pred q, %2, %3, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(out) : "r"(pred), "r"(g + 1), "r"(nt) : "memory");
    return out;
}

// Stages the ww x wh window of a pyramid level whose origin is (wx0, wy0) into `win`: one TMA
// tensor-tile load (zero fill outside the image) waited on by the whole group, or cooperative
// 16-byte loads for boxes wider than a TMA tile.  Returns false if the TMA wait timed out.
__device__ __forceinline__ bool stage_window(const NormalsArgs& A, const CUtensorMap* tmap, bool by_tma, const uint8_t* img,
                                             const fm3d_level& lv, uint8_t* win, int ww, int wh, int wx0, int wy0,
                                             uint64_t* bar, int* tma_phase, int groups, int g, int NT, int tid) {
    if (ww <= 0 || wh <= 0) return false;
    if (by_tma) {
        const uint32_t parity = (uint32_t)*tma_phase;
        if (tid == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, (uint32_t)(ww * wh));
            tma_load_2d(win, tmap, wx0, wy0, bar);
        }
        bool ok = false;
        for (int spin = 0; spin < (1 << 22); spin++) {
            if (mbar_try_wait(bar, parity)) { ok = true; break; }
        }
        const bool staged = gsync_and(groups, g, NT, ok ? 1 : 0) != 0;
        if (tid == 0) {
            *tma_phase ^= 1;
            if (!staged) atomicExch(A.error_flag, 1);
        }
        return staged;
    }
    const int wq = ww >> 4;
    for (int i = tid; i < wq * wh; i += NT) {
        const int yy = i / wq, xq = i - yy * wq;
        const int gx = wx0 + 16 * xq, gy = wy0 + yy;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (gy >= 0 && gy < lv.h && gx >= 0 && gx + 16 <= lv.pitch) {
            v = *reinterpret_cast<const uint4*>(img + (size_t)gy * lv.pitch + gx);
        } else if (gy >= 0 && gy < lv.h) {
            __align__(16) uint8_t b[16];
            for (int k = 0; k < 16; k++) b[k] = (gx + k >= 0 && gx + k < lv.w) ? img[(size_t)gy * lv.pitch + gx + k] : 0;
            v = *reinterpret_cast<uint4*>(b);
        }
        *reinterpret_cast<uint4*>(win + (size_t)yy * ww + 16 * xq) = v;
    }
    gsync(groups, g, NT);
    return true;
}

// What a feature needs before its first level, by all threads of its group: the disc lattice of extractPixelsContour as a row
// table, the per-feature constants (thread 0) and the ideal rays of all disc pixels as offsets from the ray through P.
__device__ FM3D_SETUP_INLINE FeatureLocals feature_prologue(const NormalsArgs& A, GroupCtl& G, float2* rays, const int f, const int tid,
                                                          const int NT, const int groups, const int g) {
    const fm3d_cam& cam = A.cam;
    const int r = A.r, W = A.pyr.lv[0].w, H = A.pyr.lv[0].h, levels = A.pyr.levels;
    RowTable* rows = &G.rows;
    FastShared* S = &G.S;
    const long long t_f0 = clock64();
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
    gsync(groups, g, NT);
    if (tid == 0) {
        int acc = 0, first = -1, last = -1;
        rows->start[0] = 0;
        for (int jr = 0; jr < 2 * r + 1; jr++) {
            const int c = rows->start[jr + 1];
            if (c > 0) { if (first < 0) first = jr; last = jr; }
            acc += c;
            rows->start[jr + 1] = acc;
        }
        rows->nrows = 2 * r + 1;
        S->first_row = first; S->last_row = last;
        S->m = acc;
        S->status = acc > 0 ? FM3D_FEAT_OK : FM3D_FEAT_NO_PIXELS;
        // a non-finite centre passes every `p < 0 || p >= size` test of the reference's lattice loop,
        // the feature then dies in its first evaluation with a NaN plane point
        if (cu != cu || cv != cv) S->status = FM3D_FEAT_ABORT_NAN;
        S->alive = acc > 0;
        S->npenalty = 0;
        for (int k = 0; k < 16; k++) S->stats[k] = 0;
        S->P[0] = Px; S->P[1] = Py; S->P[2] = Pz;
        const double nrm = sqrt(Px * Px + Py * Py + Pz * Pz);
        S->normal[0] = Px / nrm; S->normal[1] = Py / nrm; S->normal[2] = Pz / nrm;  // (:343)
        if (A.nfev && A.mode == 0) for (int l = 0; l <= levels; l++) A.nfev[(size_t)f * (levels + 1) + l] = 0;
        // centre ray, and the camera-2 projection of P: P lies on every candidate plane, so this
        // is where the centre of the disc lands whatever the normal (reference point of the
        // offset form)
        // reference ray of the offsets: the ray through P itself (not the 5-iteration undistort of the centre pixel, which
        // misses it by the residual of that iteration; every disc pixel still gets its own 5-iteration ray below)
        S->vc[0] = Px / Pz; S->vc[1] = Py / Pz;
        const double X2 = cam.R[0] * Px + cam.R[1] * Py + cam.R[2] * Pz + cam.t[0];
        const double Y2 = cam.R[3] * Px + cam.R[4] * Py + cam.R[5] * Pz + cam.t[1];
        const double Z2 = cam.R[6] * Px + cam.R[7] * Py + cam.R[8] * Pz + cam.t[2];
        const double iz = Z2 != 0.0 ? 1.0 / Z2 : 1.0;
        S->xc = X2 * iz; S->yc = Y2 * iz;
    }
    gsync(groups, g, NT);
    const int m = S->m;
    if (A.m_out && tid == 0) A.m_out[f] = m;
    const double vcx = S->vc[0], vcy = S->vc[1];
    const float vcxf = (float)vcx, vcyf = (float)vcy;

    // ideal rays of all disc pixels (normal-independent), as offsets from the centre ray
    if (m > 0) {
        int row = 0;
        for (int idx = tid; idx < m; idx += NT) {
            while (idx >= rows->start[row + 1]) row++;
            const double px = cu + (double)(rows->ilo[row] + (idx - rows->start[row]));
            const double py = cv + (double)rows->jrow[row];
            double vx, vy;
            undistort_ray(cam, px, py, vx, vy);
            ray_set(rays, idx, (float)(vx - vcx), (float)(vy - vcy));
        }
    }
    if (tid == 0) S->stats[14] += (unsigned long long)(clock64() - t_f0);   // prologue: lattice + rays
    FeatureLocals F;
    F.cu = cu; F.cv = cv; F.m = m; F.vcxf = vcxf; F.vcyf = vcyf;
    return F;
}

// What a pyramid level needs before its first pass, by all threads of the group: window origins (thread 0), the image-1
// samples of the level (from a staged image-1 window), the image-2 window, the level constants L, and on warp 0 the start of
// the level's LM (or, modes 1 / 2, the first evaluation point).  Returns the image-1 gate flags of this thread's pixels.
template <bool ncc>
__device__ FM3D_SETUP_INLINE unsigned level_setup(const NormalsArgs& A, GroupCtl& G, uint8_t* win, float* i1, const int f, const int lvl,
                                                const FeatureLocals& F, const int tid, const int NT, const int groups, const int g,
                                                LevelConst& L_out, const bool final_sync = true) {
    const fm3d_cam& cam = A.cam;
    const int r = A.r;
    const double cu = F.cu, cv = F.cv;
    const int m = F.m;
    RowTable* rows = &G.rows;
    double* red = G.red;
    FastShared* S = &G.S;
    FastPass* PP = &G.PP;
    uint64_t* bar = &G.bar;
    const int lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    const long long t_l0 = clock64();
    const fm3d_level lv = A.pyr.lv[lvl];
    const double scale = 1.0 / (double)(1 << lvl);      // actual_scale_ (:226-241)
    const double inv_scale = 1.0 / scale;
    const uint8_t* img1 = A.pyr.base[0] + lv.off;
    LevelConst L;
    L.img2 = A.pyr.base[1] + lv.off;
    L.w = lv.w; L.h = lv.h; L.pitch = lv.pitch;
    L.win = win; L.ww = (unsigned)A.win_w[lvl];
    const int wh = A.win_h[lvl];
    L.cols = (float)lv.w; L.rows = (float)lv.h;     // scale * (cols_l / scale)
    L.xc = (float)S->xc; L.yc = (float)S->yc;
    L.c0 = 0.0f;

    // window origins.  Image 2: centred on the projection of P (P lies on every candidate plane).
    // Image 1: the disc itself, scaled.  TMA needs the box start address 16-byte aligned: x
    // origins are multiples of 16 pixels.
    if (tid == 0) {
        double u2, v2;
        fm3d_distort_K<double>(S->xc, S->yc, cam.k1, cam.k2, cam.p1, cam.p2, cam.k3, cam.fx, cam.fy, cam.cx, cam.cy, u2, v2);
        double wxc = floor(scale * u2), wyc = floor(scale * v2);
        if (!(wxc > -1e6 && wxc < 1e6)) wxc = 0;
        if (!(wyc > -1e6 && wyc < 1e6)) wyc = 0;
        S->wx0 = (((int)wxc - (int)L.ww / 2 + 8) >> 4) << 4;
        S->wy0 = (int)wyc - wh / 2;
        double w1x = floor(scale * (cu - (double)r)) - 1.0, w1y = floor(scale * (cv - (double)r)) - 1.0;
        if (!(w1x > -1e6 && w1x < 1e6)) w1x = 0;
        if (!(w1y > -1e6 && w1y < 1e6)) w1y = 0;
        S->w1x0 = ((int)w1x >> 4) << 4;
        S->w1y0 = (int)w1y;
    }
    gsync(groups, g, NT);  // also: everybody is done with the previous level's window
    L.wx0 = S->wx0; L.wy0 = S->wy0;
    const bool by_tma = A.use_tma && A.win_tma[lvl];

    // image-1 intensities of the level (updateImage1PixelsIntensity, :576-589), sampled from the
    // staged image-1 window; a pixel whose taps leave window or image takes the global path
    unsigned lvl_flags = 0;
    {
        const int w1x0 = S->w1x0, w1y0 = S->w1y0;
        const bool staged1 = stage_window(A, &A.tmap1[lvl], by_tma, img1, lv, win, (int)L.ww, wh, w1x0, w1y0, bar,
                                          &S->tma_phase, groups, g, NT, tid);
        const int x_lo = max(w1x0, 0), x_hi = min(w1x0 + (int)L.ww, lv.w) - 1;   // x0 in [x_lo, x_hi)
        const int y_lo = max(w1y0, 0), y_hi = min(w1y0 + wh, lv.h) - 1;
        const int x_cnt = staged1 ? max(x_hi - x_lo, 0) : 0, y_cnt = staged1 ? max(y_hi - y_lo, 0) : 0;
        // one warp per row of the disc: what depends on the row only (y tap, y weight, the row's part of the isPixelGood gate,
        // the tap row in the window) is formed once per row, the samples of a row are stored side by side
        double i1_part = 0.0;
        const int n_rows = rows->nrows;
        for (int row = wid; row < n_rows; row += NW) {
            const int s0 = rows->start[row], cnt = rows->start[row + 1] - s0;
            if (cnt <= 0) continue;
            const int ilo = rows->ilo[row];
            const double py = cv + (double)rows->jrow[row];
            const bool y_bad = (py < 0) || (py > inv_scale * lv.h);
            const float sy = (float)(scale * py);
            const float fy0 = floorf(sy);
            const int y0 = (int)fy0;
            const float ay = __fsub_rn(sy, fy0);
            const bool y_in = (unsigned)(y0 - y_lo) < (unsigned)y_cnt;
            const uint8_t* wrow = win + (y0 - w1y0) * (int)L.ww - w1x0;
            for (int i = lane; i < cnt; i += 32) {
                const double px = cu + (double)(ilo + i);
                if (y_bad || (px < 0) || (px > inv_scale * lv.w)) lvl_flags |= FLAG_PIX;   // fm3d_pixel_good
                const float sx = (float)(scale * px);
                const float fx0 = floorf(sx);
                const int x0 = (int)fx0;
                float v;
                if (y_in && (unsigned)(x0 - x_lo) < (unsigned)x_cnt) {
                    const uint8_t* pw = wrow + x0;
                    v = fm3d_lerp4(fm3d_u8f(pw[0]), fm3d_u8f(pw[L.ww]), fm3d_u8f(pw[1]), fm3d_u8f(pw[L.ww + 1]),
                                   __fsub_rn(sx, fx0), ay);
                } else {
                    v = fm3d_bilinear_global(img1, lv.w, lv.h, lv.pitch, sx, sy);
                }
                i1[s0 + i] = v;
                i1_part += (double)v;
            }
        }
        if (ncc) {
            // cost_mode NCC: the image-1 samples are kept CENTRED (u = I1 - c0, c0 = their mean rounded to float) together
            // with sum u and sum u^2; image-2 samples are centred with the same constant before they are summed, so that
            // the variances below are differences of small numbers' squares (fp64 sums of fp32 products)
            const double ws = warp_sum(i1_part);
            if (lane == 0) red[wid * NSUM] = ws;
            gsync(groups, g, NT);
            double tot = 0.0;
            for (int w = 0; w < NW; w++) tot += red[w * NSUM];     // same order in every thread: one value for the group
            const float c0 = (float)(tot / (double)m);
            L.c0 = c0;
            double pu = 0.0, puu = 0.0;
            for (int idx = tid; idx < m; idx += NT) {
                const float a = __fsub_rn(i1[idx], c0);
                i1[idx] = a;
                pu += (double)a;
                puu = fma((double)a, (double)a, puu);
            }
            const double wu = warp_sum(pu), wuu = warp_sum(puu);
            gsync(groups, g, NT);                                  // every thread has read the first totals
            if (lane == 0) { red[wid * NSUM] = wu; red[wid * NSUM + 1] = wuu; }
            gsync(groups, g, NT);
            if (tid == 0) {
                double a = 0.0, b = 0.0;
                for (int w = 0; w < NW; w++) { a += red[w * NSUM]; b += red[w * NSUM + 1]; }
                S->ncc_su = a; S->ncc_suu = b;
            }
        }
        gsync(groups, g, NT);   // everybody is done with the image-1 window
    }

    // image-2 window
    const bool staged = stage_window(A, &A.tmap[lvl], by_tma, L.img2, lv, win, (int)L.ww, wh, L.wx0, L.wy0, bar,
                                     &S->tma_phase, groups, g, NT, tid);
    // taps (x0,y0),(x0+1,y0+1) must lie inside the window and inside the image
    {
        int x_lo = max(L.wx0, 0), x_hi = min(L.wx0 + (int)L.ww, lv.w) - 1;  // x0 in [x_lo, x_hi)
        int y_lo = max(L.wy0, 0), y_hi = min(L.wy0 + wh, lv.h) - 1;
        L.lx_min = x_lo - L.wx0; L.lx_cnt = staged ? max(x_hi - x_lo, 0) : 0;
        L.ly_min = y_lo - L.wy0; L.ly_cnt = staged ? max(y_hi - y_lo, 0) : 0;
    }
    L.coff = (FLOOR_MAGIC_BITS + (unsigned)L.wy0) * L.ww + FLOOR_MAGIC_BITS + (unsigned)L.wx0;
    L.amax = L.ww * (unsigned)wh - L.ww - 2u;
    L.koff = smem_u32(win) - L.coff;
    L.ahi = smem_u32(win) + L.amax;

    // warp 0: start the LM of this level (optimize(), :247-292)
    if (wid == 0) {
        double phi, theta;
        if (A.mode == 0) {
            const double* nv = S->normal;
            theta = atan2(nv[2], sqrt(nv[0] * nv[0] + nv[1] * nv[1]));  // car2sph (tools.cpp:767-771)
            phi = atan2(nv[1], nv[0]);
            if (lane == 0) {
                PP->slow = 0;
                fm3d_lm2_init(&S->lm, phi, theta, A.eps_lmmin, A.patience);
                S->where = AT_X_JAC;
                S->base_valid = 0; S->jbase_valid = 0; S->tnext = 0; S->tvalid[0] = S->tvalid[1] = S->tvalid[2] = S->tvalid[3] = 0;
                S->fuse_hint = 1;
            }
            publish_pass(PP, S, cam, phi, theta, PASS_JAC, A.penalty_mode, sqrt(fmax(A.eps_lmmin, FM3D_DBL_EPS)), lane);
        } else {
            if (A.phi_theta) {
                phi = A.phi_theta[2 * f]; theta = A.phi_theta[2 * f + 1];
            } else {   // the initial normal of the optimiser: the viewing ray (normaloptimizer.cpp:343)
                const double* nv = S->normal;
                theta = atan2(nv[2], sqrt(nv[0] * nv[0] + nv[1] * nv[1]));
                phi = atan2(nv[1], nv[0]);
            }
            if (lane == 0) {
                PP->slow = 0;
                S->lm.nfev = 0;
                S->lm.eps = 1e-5;
                S->where = AT_XT_PLAIN;
                S->base_valid = 0; S->jbase_valid = 0; S->tnext = 0; S->tvalid[0] = S->tvalid[1] = S->tvalid[2] = S->tvalid[3] = 0;
                S->cand = 0;
                S->sweep_c[0] = phi; S->sweep_c[1] = theta;
                S->best_cost = __longlong_as_double(0x7ff0000000000000LL);
                S->best_idx = -1;
            }
            if (A.mode == 2) {   // first candidate of the grid
                phi -= 0.5 * (double)(A.sweep_nphi - 1) * A.sweep_dphi;
                theta -= 0.5 * (double)(A.sweep_ntheta - 1) * A.sweep_dtheta;
            }
            publish_pass(PP, S, cam, phi, theta, PASS_VALUE, A.penalty_mode, 1e-5, lane);
        }
    }
    // two-slot kernel with normals_level_sync = 0: the shared warps do not wait for the LM warp's start of the level -- the next
    // turn of the slot opens with the slot's barrier, after the LM warp has announced it (slot_ready)
    if (final_sync) gsync(groups, g, NT);

    if (tid == 0) S->stats[15] += (unsigned long long)(clock64() - t_l0);   // level set-up: window + image-1 samples
    L_out = L;
    return lvl_flags;
}

// RAYS_SMEM / I1_SMEM: where the per-pixel ray offsets (8 B) and image-1 samples (4 B) of the
// feature live: shared memory, or a per-group scratch in global memory that stays L2-resident
// (one CTA streams it once per pass: 12 B x 12 853 pixels at r = 64).
// NCC_COST: cost_mode NCC is a separate instantiation so that the reference's SSD kernel keeps its register allocation.
template <bool RAYS_SMEM, bool I1_SMEM, bool NCC_COST>
__global__ void __launch_bounds__(FAST_NT, 1)
normals_fast_kernel(const __grid_constant__ NormalsArgs A) {
    extern __shared__ __align__(128) uint8_t smem_all[];
    const int groups = A.groups;
    const int NT = blockDim.x / groups;
    const int g = threadIdx.x / NT;
    const int tid = threadIdx.x - g * NT;
    uint8_t* smem = smem_all + (size_t)g * A.group_smem;
    uint8_t* win = smem;
    float2* rays;
    float* i1;
    uint8_t* tail = smem + A.win_bytes;
    const size_t slot = (size_t)blockIdx.x * groups + g;
    if (RAYS_SMEM) {
        rays = reinterpret_cast<float2*>(tail);
        tail += sizeof(float2) * (size_t)A.mcap;
    } else {
        rays = A.rays_g + slot * A.mcap;
    }
    if (I1_SMEM) {
        i1 = reinterpret_cast<float*>(tail);
        tail += sizeof(float) * (size_t)A.mcap;
    } else {
        i1 = A.i1_g + slot * A.mcap;
    }
    // per-group control state in STATIC shared memory: the compiler then knows the address space
    // (LDS/STS with immediate offsets instead of generic loads with 64-bit address arithmetic), which
    // matters most in the single-lane LM step
    __shared__ GroupCtl ctl[MAX_GROUPS];
    GroupCtl& G = ctl[g];
    RowTable* rows = &G.rows;
    double* red = G.red;            // [16 warps][6]
    unsigned* wflags = G.wflags;    // [16 warps]
    FastShared* S = &G.S;
    FastPass* PP = &G.PP;
    uint64_t* bar = &G.bar;

    const int lane = tid & 31, wid = tid >> 5, NW = NT >> 5;
    constexpr bool ncc = NCC_COST;
    const fm3d_cam& cam = A.cam;
    const int r = A.r, W = A.pyr.lv[0].w, H = A.pyr.lv[0].h, levels = A.pyr.levels;
    const float cmax = (float)(int)(2 * cam.zmax);  // int cMax = 2*z_threshold_max_ (:648)

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        S->tma_phase = 0;
    }
    __syncthreads();

    for (;;) {
        // ------------------------------------------------------------ fetch a feature
        if (tid == 0) S->feature = atomicAdd(A.work_counter, 1);
        gsync(groups, g, NT);
        const int f = S->feature;
        if (f >= A.n) break;
        const FeatureLocals F = feature_prologue(A, G, rays, f, tid, NT, groups, g);
        const double Px = A.xyz[3 * f], Py = A.xyz[3 * f + 1], Pz = A.xyz[3 * f + 2];
        const double cu = F.cu, cv = F.cv;
        const int m = F.m;
        const float vcxf = F.vcxf, vcyf = F.vcyf;
        // boundary lattice: both ends of every row, and the whole first and last row
        const int nrows = 2 * r + 1;
        const int first_row = S->first_row, last_row = S->last_row;
        const int n_first = m > 0 ? rows->start[first_row + 1] - rows->start[first_row] : 0;
        const int n_last = m > 0 ? rows->start[last_row + 1] - rows->start[last_row] : 0;
        const int n_boundary = m > 0 ? 2 * nrows + n_first + n_last : 0;

        // ------------------------------------------------------------ coarse-to-fine LM
        const int lvl_hi = A.mode == 0 ? levels : A.eval_level;
        const int lvl_lo = A.mode == 0 ? 0 : A.eval_level;
        bool alive = m > 0;
        for (int lvl = lvl_hi; lvl >= lvl_lo && alive; lvl--) {
            LevelConst L;
            unsigned lvl_flags = level_setup<ncc>(A, G, win, i1, f, lvl, F, tid, NT, groups, g, L);
            // -------------------------------------------------------- mode 2: the candidate grid in batches
            if (A.mode == 2 && A.sweep_batch > 1) {
                const int K = A.sweep_nphi * A.sweep_ntheta;
                bool fall_back = false;         // a tap left the staged window: the one-by-one loop takes over
                for (int c0 = 0; c0 < K && !fall_back; c0 += SWEEP_B) {
                    const int nb = min(SWEEP_B, K - c0);
                    const long long t_a = clock64();
                    if (wid == 0) {
                        for (int k = 0; k < nb; k++) {
                            const int c = c0 + k, ip = c / A.sweep_ntheta, it = c - ip * A.sweep_ntheta;
                            const double phi = S->sweep_c[0] + ((double)ip - 0.5 * (double)(A.sweep_nphi - 1)) * A.sweep_dphi;
                            const double theta = S->sweep_c[1] + ((double)it - 0.5 * (double)(A.sweep_ntheta - 1)) * A.sweep_dtheta;
                            if (lane == 0) G.PPk[k].slow = 0;
                            publish_pass(&G.PPk[k], S, cam, phi, theta, PASS_VALUE, A.penalty_mode, 1e-5, lane);
                            __syncwarp();
                            if (lane == 0) G.wk[k] = S->w[0];
                            __syncwarp();
                        }
                    }
                    gsync(groups, g, NT);
                    // a NaN candidate normal ends the feature (normaloptimizer.cpp:81-85): evaluate the ones before it
                    int ne = nb;
                    for (int k = nb - 1; k >= 0; k--) if (G.PPk[k].kind == PASS_STOP) ne = k;
                    unsigned fl[SWEEP_B];
#pragma unroll
                    for (int k = 0; k < SWEEP_B; k++) fl[k] = 0u;
                    fl[0] = lvl_flags;          // image-1 gate of the level: charged to the first evaluation, as below
                    lvl_flags = 0;
                    for (int k = tid; k < n_boundary; k += NT) {
                        int idx;
                        if (k < 2 * nrows) {
                            const int row = k >> 1;
                            const int s0 = rows->start[row], s1 = rows->start[row + 1];
                            if (s1 <= s0) continue;
                            idx = (k & 1) ? s1 - 1 : s0;
                        } else if (k < 2 * nrows + n_first) {
                            idx = rows->start[first_row] + (k - 2 * nrows);
                        } else {
                            idx = rows->start[last_row] + (k - 2 * nrows - n_first);
                        }
                        const float2 ray = ray_at(rays, idx);
#pragma unroll
                        for (int kk = 0; kk < SWEEP_B; kk++)
                            if (kk < ne) fl[kk] |= boundary_flags(A.fc, lvl, G.PPk[kk], L, vcxf, vcyf, cmax, ray);
                    }
                    Acc acc[SWEEP_B];
#pragma unroll
                    for (int k = 0; k < SWEEP_B; k++) {
                        acc[k].s0 = acc[k].t1 = acc[k].t2 = 0.0;
#pragma unroll
                        for (int j = 0; j < 9; j++) acc[k].f[j] = 0.0f;
                    }
                    if (ne == SWEEP_B) run_pixels_multi<SWEEP_B, !RAYS_SMEM>(A.fc, lvl, G.PPk, L, rays, i1, m, tid, NT, acc);
                    else if (ne == 3) run_pixels_multi<3, !RAYS_SMEM>(A.fc, lvl, G.PPk, L, rays, i1, m, tid, NT, acc);
                    else if (ne == 2) run_pixels_multi<2, !RAYS_SMEM>(A.fc, lvl, G.PPk, L, rays, i1, m, tid, NT, acc);
                    else if (ne == 1) run_pixels_multi<1, !RAYS_SMEM>(A.fc, lvl, G.PPk, L, rays, i1, m, tid, NT, acc);
#pragma unroll
                    for (int k = 0; k < SWEEP_B; k++) {
                        const double a = warp_sum(acc[k].s0);
                        const unsigned f_or = __reduce_or_sync(0xffffffffu, fl[k]);
                        if (lane == 0) { red[wid * NSUM + k] = a; G.wflagsk[wid * SWEEP_B + k] = f_or; }
                    }
                    const long long t_b0 = clock64();
                    gsync(groups, g, NT);
                    const long long t_b = clock64();
                    if (wid == 0) {
                        double sk[SWEEP_B];
                        unsigned fk[SWEEP_B];
#pragma unroll
                        for (int k = 0; k < SWEEP_B; k++) {
                            double v = lane < NW ? red[lane * NSUM + k] : 0.0;
#pragma unroll
                            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                            sk[k] = v;
                            fk[k] = __reduce_or_sync(0xffffffffu, lane < NW ? G.wflagsk[lane * SWEEP_B + k] : 0u);
                        }
                        if (lane == 0) {
                            int resume = -1;
                            for (int k = 0; k < ne && resume < 0; k++)
                                if ((fk[k] & FLAG_WINDOW) && !(fk[k] & 7)) resume = c0 + k;
                            const int n_rec = resume >= 0 ? resume - c0 : ne;
                            for (int k = 0; k < n_rec; k++) {
                                const int c = c0 + k;
                                const bool bad = (fk[k] & 7) || sk[k] != sk[k];
                                const double cval = bad ? __longlong_as_double(0x7ff8000000000000LL) : G.wk[k] * G.wk[k] * sk[k];
                                if (A.cost) A.cost[(size_t)f * K + c] = cval;
                                if (!bad && cval < S->best_cost) { S->best_cost = cval; S->best_idx = c; }
                            }
                            S->stats[0] += (unsigned long long)n_rec;
                            S->stats[5] += (unsigned long long)m * (unsigned long long)ne;
                            S->lm.nfev += n_rec;
                            G.resume_at = resume;
                            S->cand = resume >= 0 ? resume : c0 + n_rec;
                            const long long t_c = clock64();
                            S->stats[8] += (unsigned long long)(t_b0 - t_a);
                            S->stats[9] += (unsigned long long)(t_b - t_b0);
                            S->stats[10] += (unsigned long long)(t_c - t_b);
                        }
                    }
                    gsync(groups, g, NT);
                    if (G.resume_at >= 0) fall_back = true;
                    if (ne < nb) break;         // NaN candidate: S->status / S->alive were set by publish_pass
                }
                if (wid == 0) {
                    if (fall_back && S->alive) {
                        // candidate S->cand again, one by one from here on (taps from global memory when needed)
                        const int c = S->cand, ip = c / A.sweep_ntheta, it = c - ip * A.sweep_ntheta;
                        const double phi = S->sweep_c[0] + ((double)ip - 0.5 * (double)(A.sweep_nphi - 1)) * A.sweep_dphi;
                        const double theta = S->sweep_c[1] + ((double)it - 0.5 * (double)(A.sweep_ntheta - 1)) * A.sweep_dtheta;
                        publish_pass(PP, S, cam, phi, theta, PASS_VALUE, A.penalty_mode, 1e-5, lane);
                    } else if (lane == 0) {
                        PP->kind = PASS_STOP;
                    }
                }
                gsync(groups, g, NT);
            }
            // -------------------------------------------------------- pass loop (two barriers per pass)
            for (;;) {
                const FastPass P = *PP;
                if (P.kind == PASS_STOP) break;
                const long long t_a = clock64();

                unsigned flags = lvl_flags;
                lvl_flags = 0;
                for (int k = tid; k < n_boundary; k += NT) {
                    int idx;
                    if (k < 2 * nrows) {
                        const int row = k >> 1;
                        const int s0 = rows->start[row], s1 = rows->start[row + 1];
                        if (s1 <= s0) continue;
                        idx = (k & 1) ? s1 - 1 : s0;
                    } else if (k < 2 * nrows + n_first) {
                        idx = rows->start[first_row] + (k - 2 * nrows);
                    } else {
                        idx = rows->start[last_row] + (k - 2 * nrows - n_first);
                    }
                    flags |= boundary_flags(A.fc, lvl, P, L, vcxf, vcyf, cmax, ray_at(rays, idx));
                }

                Acc acc;
                acc.s0 = acc.t1 = acc.t2 = 0.0;
#pragma unroll
                for (int k = 0; k < 9; k++) acc.f[k] = 0.0f;
                if (!ncc) {
                    if (P.kind == PASS_JAC) {
                        if (!P.slow) run_pixels<true, false, !RAYS_SMEM>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                        else run_pixels<true, true, false>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                    } else {
                        if (!P.slow) run_pixels<false, false, !RAYS_SMEM>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                        else run_pixels<false, true, false>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                    }
                } else {
                    if (P.kind == PASS_JAC) {
                        if (!P.slow) run_pixels<true, false, !RAYS_SMEM, true>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                        else run_pixels<true, true, false, true>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                    } else {
                        if (!P.slow) run_pixels<false, false, !RAYS_SMEM, true>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                        else run_pixels<false, true, false, true>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                    }
                }
                // per-warp sums: SSD 1 (+5 with the Jacobian), NCC 3 (+9)
                {
                    double* rw = red + wid * NSUM;
                    const double a0 = warp_sum(acc.s0);
                    if (lane == 0) rw[0] = a0;
                    if (ncc) {
                        const double a1 = warp_sum(acc.t1), a2 = warp_sum(acc.t2);
                        if (lane == 0) { rw[1] = a1; rw[2] = a2; }
                    }
                    const int nj = P.kind == PASS_JAC ? (ncc ? 9 : 5) : 0, jo = ncc ? 3 : 1;
                    for (int k = 0; k < 9; k++) {
                        if (k >= nj) break;
                        const double a = warp_sum((double)acc.f[k]);
                        if (lane == 0) rw[jo + k] = a;
                    }
                    if (lane == 0 && P.kind != PASS_JAC && !ncc) { rw[1] = rw[2] = rw[3] = rw[4] = rw[5] = 0.0; }
                }
                flags = __reduce_or_sync(0xffffffffu, flags);
                if (lane == 0) wflags[wid] = flags;
                const long long t_b0 = clock64();
                gsync(groups, g, NT);
                const long long t_b = clock64();

                if (wid == 0) lm_advance<ncc>(A, G, P, m, f, NW, lane, t_a, t_b0, t_b);
                gsync(groups, g, NT);
            }
            alive = S->alive != 0;

            if (tid == 0 && A.mode == 0) {
                const fm3d_lm2& lm = S->lm;
                if (alive) {
                    // sph2car of the solution (:289)
                    S->normal[0] = cos(lm.x[1]) * cos(lm.x[0]);
                    S->normal[1] = cos(lm.x[1]) * sin(lm.x[0]);
                    S->normal[2] = sin(lm.x[1]);
                    if (A.cost) A.cost[f] = lm.ff;
                }
                if (A.nfev) A.nfev[(size_t)f * (levels + 1) + lvl] = lm.nfev;
            }
            gsync(groups, g, NT);
        }

        // ------------------------------------------------------------ epilogue
        if (tid < 16 && A.stats) {
            const unsigned long long v = tid == 7 ? 1ull : S->stats[tid];
            if (v) atomicAdd(A.stats + tid, v);
        }
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
            } else if (A.mode == 2) {
                if (A.best_idx) A.best_idx[f] = st == FM3D_FEAT_OK ? S->best_idx : -1;
                if (A.best_cost) A.best_cost[f] = S->best_cost;
            } else if (st != FM3D_FEAT_OK) {
                A.cost[f] = __longlong_as_double(0x7ff8000000000000LL);
            }
        }
        gsync(groups, g, NT);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two features per group of eight warps, taking turns ("ping-pong"): the serial LM step of one feature runs on its LM
// warp while the other seven warps evaluate the pass of the partner feature, so the step between two passes leaves the
// critical path (in normals_fast_kernel the four warps of a group wait for it: a quarter of all warp time, ncu r02l).
//   * a CTA has two super-groups of eight warps; super-group sg owns slots 2 sg and 2 sg + 1 (window, scratch, LM state:
//     what a group of normals_fast_kernel owns);
//   * slot s is served by seven warps: all but the LM warp of the other slot.  With the warps relabelled wr = w xor lm_s the
//     slot's LM warp is wr = 0 and the excluded warp wr = lm_0 xor lm_1 in both slots, so pixel partition, summation order
//     and therefore the results are the same in both slots (and a feature stays in its slot);
//   * a warp serves slot 0, then slot 1, then slot 0 ...; a turn is one STEP of the slot: fetch + prologue + level set-up,
//     level set-up, or one pass.  A turn starts with the slot's barrier (224 threads): it opens when the slot's LM warp has
//     published the next pass.  After a pass the six shared warps only ARRIVE at the slot's second barrier and move on to
//     the other slot; the LM warp waits there for the sums and advances the LM meanwhile.
// Mode 0 (optimise) with the four-window layout only; everything else runs in normals_fast_kernel.
enum { STEP_FETCH = 0, STEP_LEVEL = 1, STEP_PASS = 2 };

__device__ __forceinline__ void nbar_arrive(int id, int nt) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nt) : "memory"); }
__device__ __forceinline__ void nbar_sync(int id, int nt) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nt) : "memory"); }

__device__ __forceinline__ bool nbar_red_or(int id, int nt, bool pred) {
    int out;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %1, 0;\n\t"
        "bar.red.or.pred q, %2, %3, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(out) : "r"(pred ? 1 : 0), "r"(id), "r"(nt) : "memory");
    return out != 0;
}
// The slot's LM warp has decided the slot's next step (everything it wrote before is visible to who sees the new count).
__device__ __forceinline__ void slot_ready(GroupCtl& G, const int lane) {
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        *(volatile int*)&G.ready_gen = G.ready_gen + 1;
    }
}
// End of a feature, on its LM warp: counters, status, normal (the epilogue of normals_fast_kernel).
__device__ __forceinline__ void finish_feature(const NormalsArgs& A, FastShared* S, const int f, const int lane) {
    if (lane < 16 && A.stats) {
        const unsigned long long v = lane == 7 ? 1ull : S->stats[lane];
        if (v) atomicAdd(A.stats + lane, v);
    }
    if (lane == 0) {
        const int st = S->status;
        A.status[f] = st;
        if (st == FM3D_FEAT_OK) {
            A.normals[3 * f] = S->normal[0]; A.normals[3 * f + 1] = S->normal[1]; A.normals[3 * f + 2] = S->normal[2];
        } else {
            const double Px = S->P[0], Py = S->P[1], Pz = S->P[2];
            const double nrm = sqrt(Px * Px + Py * Py + Pz * Pz);
            A.normals[3 * f] = Px / nrm; A.normals[3 * f + 1] = Py / nrm; A.normals[3 * f + 2] = Pz / nrm;
            if (A.cost) A.cost[f] = __longlong_as_double(0x7ff8000000000000LL);
        }
        if (A.npenalty) A.npenalty[f] = S->npenalty;
    }
}

template <bool NCC_COST>
__global__ void __launch_bounds__(FAST_NT, 1)
normals_pp_kernel(const __grid_constant__ NormalsArgs A) {
    extern __shared__ __align__(128) uint8_t smem_all[];
    __shared__ GroupCtl ctl[MAX_GROUPS];
    constexpr bool ncc = NCC_COST;
    constexpr int NT = 224, NW = 7;             // threads / warps that serve a slot
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sg = warp >> 3, w8 = warp & 7;
    // LM warps of the two slots of a super-group: warps 0, 5 / 10, 15 of the CTA, one per scheduler (warp % 4); the labels
    // wr = w8 xor lm differ by lm0 xor lm1 = 5 in both slots
    const int lm0 = 2 * sg, lm1 = lm0 ^ 5;
    const fm3d_cam& cam = A.cam;
    const int r = A.r, levels = A.pyr.levels;
    const float cmax = (float)(int)(2 * cam.zmax);  // int cMax = 2*z_threshold_max_ (:648)
    const int nrows = 2 * r + 1;

    if (threadIdx.x < MAX_GROUPS) {
        GroupCtl& G0 = ctl[threadIdx.x];
        mbar_init(&G0.bar, 1);
        fence_mbar_init();
        G0.S.tma_phase = 0;
        G0.kind = STEP_FETCH;
        G0.ready_gen = 1;
    }
    __syncthreads();

    // Turn order.  The LM warp of a slot serves its own slot only.  The six shared warps serve whichever slot is ready,
    // preferring the one they did not serve last: a slot is ready when its LM warp has made more steps ready (ready_gen) than
    // the shared warps have started (served).  They agree on it with a barrier reduction over their 192 threads (ready_gen
    // only grows, so "one of us saw it ready" is a safe verdict for all).
    bool done0 = false, done1 = false;
    int served0 = 0, served1 = 0, last = 1;
    const int BAR_TURN = 1 + 2 * MAX_GROUPS + sg;   // named barrier 9 + sg
    for (;;) {
        int s;
        if (w8 == lm0) { if (done0) break; s = 0; }
        else if (w8 == lm1) { if (done1) break; s = 1; }
        else {
            if (done0 && done1) break;
            const int pref = done0 ? 1 : (done1 ? 0 : 1 - last);
            const bool other_alive = !(pref ? done0 : done1);
            for (;;) {
                const volatile int* gp = &ctl[2 * sg + pref].ready_gen;
                if (nbar_red_or(BAR_TURN, 192, *gp > (pref ? served1 : served0))) { s = pref; break; }
                if (other_alive) {
                    const volatile int* go = &ctl[2 * sg + 1 - pref].ready_gen;
                    if (nbar_red_or(BAR_TURN, 192, *go > (pref ? served0 : served1))) { s = 1 - pref; break; }
                }
                __nanosleep(256);       // neither slot is ready: leave the issue slots to the LM warps and the other super-group
            }
            last = s;
        }
        {
            if (s == 0) served0++; else served1++;
            const int wr = w8 ^ (s ? lm1 : lm0);
            const int sl = 2 * sg + s;
            GroupCtl& G = ctl[sl];
            RowTable* rows = &G.rows;
            double* red = G.red;
            unsigned* wflags = G.wflags;
            FastShared* S = &G.S;
            FastPass* PP = &G.PP;
            uint8_t* win = smem_all + (size_t)sl * A.group_smem;
            const size_t slot = (size_t)blockIdx.x * MAX_GROUPS + sl;
            float2* rays = A.rays_g + slot * A.mcap;
            float* i1 = A.i1_g + slot * A.mcap;
            const int rank = wr < 5 ? wr : wr - 1;      // 0 (LM warp) .. 6
            const int tid = rank * 32 + lane;
            const int BAR = sl;                          // gsync(2, BAR, NT): named barrier 1 + sl
            const int BAR_SUMS = 1 + MAX_GROUPS + sl;    // named barrier 5 + sl

            gsync(2, BAR, NT);                           // the slot's LM warp has decided what comes next
            int kind = G.kind;
            if (kind == STEP_FETCH) {
                if (tid == 0) S->feature = atomicAdd(A.work_counter, 1);
                gsync(2, BAR, NT);
                const int f = S->feature;
                if (f >= A.n) {                          // queue empty: the slot retires (all seven warps read the same value)
                    if (s == 0) done0 = true; else done1 = true;
                    continue;
                }
                const FeatureLocals F = feature_prologue(A, G, rays, f, tid, NT, 2, BAR);
                if (tid == 0) { G.Fsave = F; G.lvl = levels; }
                if (F.m <= 0) {                          // no pixels: nothing to optimise
                    if (rank == 0) { finish_feature(A, S, f, lane); slot_ready(G, lane); }
                    continue;                            // G.kind stays STEP_FETCH
                }
                kind = STEP_LEVEL;
                gsync(2, BAR, NT);
            }
            const int f = S->feature;
            if (kind == STEP_LEVEL) {
                const FeatureLocals F = G.Fsave;
                const int lvl = G.lvl;
                if (tid == 0) G.lvl_flags = 0u;
                LevelConst L;
                const unsigned lf = level_setup<ncc>(A, G, win, i1, f, lvl, F, tid, NT, 2, BAR, L, A.level_sync != 0);
                if (lf) atomicOr(&G.lvl_flags, lf);
                if (tid == 0) { G.Lsave = L; G.kind = STEP_PASS; }
                if (rank == 0) slot_ready(G, lane);
                continue;
            }

            // ---------------------------------------------------------------- one pass
            const FeatureLocals F = G.Fsave;
            const LevelConst L = G.Lsave;
            const int lvl = G.lvl, m = F.m;
            const FastPass P = *PP;
            const long long t_a = clock64();
            const int first_row = S->first_row, last_row = S->last_row;
            const int n_first = rows->start[first_row + 1] - rows->start[first_row];
            const int n_last = rows->start[last_row + 1] - rows->start[last_row];
            const int n_boundary = 2 * nrows + n_first + n_last;
            unsigned flags = 0u;
            if (tid == 0) { flags = G.lvl_flags; G.lvl_flags = 0u; }
            for (int k = tid; k < n_boundary; k += NT) {
                int idx;
                if (k < 2 * nrows) {
                    const int row = k >> 1;
                    const int s0 = rows->start[row], s1 = rows->start[row + 1];
                    if (s1 <= s0) continue;
                    idx = (k & 1) ? s1 - 1 : s0;
                } else if (k < 2 * nrows + n_first) {
                    idx = rows->start[first_row] + (k - 2 * nrows);
                } else {
                    idx = rows->start[last_row] + (k - 2 * nrows - n_first);
                }
                flags |= boundary_flags(A.fc, lvl, P, L, F.vcxf, F.vcyf, cmax, ray_at(rays, idx));
            }
            Acc acc;
            acc.s0 = acc.t1 = acc.t2 = 0.0;
#pragma unroll
            for (int k = 0; k < 9; k++) acc.f[k] = 0.0f;
            if (P.kind == PASS_JAC) {
                if (!P.slow) run_pixels<true, false, true, ncc>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                else run_pixels<true, true, false, ncc>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
            } else {
                if (!P.slow) run_pixels<false, false, true, ncc>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
                else run_pixels<false, true, false, ncc>(A.fc, lvl, P, L, rays, i1, m, tid, NT, acc);
            }
            {
                double* rw = red + rank * NSUM;
                const double a0 = warp_sum(acc.s0);
                if (lane == 0) rw[0] = a0;
                if (ncc) {
                    const double a1 = warp_sum(acc.t1), a2 = warp_sum(acc.t2);
                    if (lane == 0) { rw[1] = a1; rw[2] = a2; }
                }
                const int nj = P.kind == PASS_JAC ? (ncc ? 9 : 5) : 0, jo = ncc ? 3 : 1;
                for (int k = 0; k < 9; k++) {
                    if (k >= nj) break;
                    const double a = warp_sum((double)acc.f[k]);
                    if (lane == 0) rw[jo + k] = a;
                }
                if (lane == 0 && P.kind != PASS_JAC && !ncc) { rw[1] = rw[2] = rw[3] = rw[4] = rw[5] = 0.0; }
            }
            flags = __reduce_or_sync(0xffffffffu, flags);
            if (lane == 0) wflags[rank] = flags;
            if (rank != 0) {
                nbar_arrive(BAR_SUMS, NT);               // on to the other slot
                continue;
            }
            // the slot's LM warp: sums -> LM -> next pass, while the other warps serve the partner slot
            const long long t_b0 = clock64();
            nbar_sync(BAR_SUMS, NT);
            const long long t_b = clock64();
            lm_advance<ncc>(A, G, P, m, f, NW, lane, t_a, t_b0, t_b);
            __syncwarp();
            if (PP->kind == PASS_STOP) {                 // level finished, or the feature was dropped
                const bool alive = S->alive != 0;
                if (lane == 0) {
                    const fm3d_lm2& lm = S->lm;
                    if (alive) {                         // sph2car of the solution (:289)
                        S->normal[0] = cos(lm.x[1]) * cos(lm.x[0]);
                        S->normal[1] = cos(lm.x[1]) * sin(lm.x[0]);
                        S->normal[2] = sin(lm.x[1]);
                        if (A.cost) A.cost[f] = lm.ff;
                    }
                    if (A.nfev) A.nfev[(size_t)f * (levels + 1) + lvl] = lm.nfev;
                }
                __syncwarp();
                if (alive && lvl > 0) {
                    if (lane == 0) { G.lvl = lvl - 1; G.kind = STEP_LEVEL; }
                } else {
                    finish_feature(A, S, f, lane);
                    if (lane == 0) G.kind = STEP_FETCH;
                }
            }
            slot_ready(G, lane);
        }
    }
}

size_t fast_static_bytes() { return sizeof(GroupCtl) * MAX_GROUPS + 256; }   // static __shared__ of the kernel (+ slack)

}  // namespace

int run_normals_fast(fm3d_ctx* ctx, NormalsArgs& A) {
    A.cam = ctx->cam;
    A.pyr = ctx->pyr;
    A.patience = ctx->opt_lm_patience;
    A.use_tma = ctx->opt_normals_tma;
    A.fuse_trials = ctx->opt_normals_fuse;
    A.memo_trials = ctx->opt_normals_memo;
    A.level_sync = ctx->opt_normals_level_sync;
    A.sweep_batch = ctx->opt_normals_sweep_batch;
    A.cost_mode = ctx->opt_normals_cost;
    if (A.cost_mode == FM3D_COST_NCC && A.mode == 2)
        return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "the dense candidate sweep evaluates the reference's SSD cost only (normals_cost = 0)");
    A.mcap = (disc_capacity(A.r) + 31) & ~31;
    {   // camera constants of the pixel loops as floats in the kernel parameter block (constant bank -> uniform operands)
        FastConsts& c = A.fc;
        const fm3d_cam& cam = ctx->cam;
        c.k1 = (float)cam.k1; c.k2 = (float)cam.k2; c.k3 = (float)cam.k3; c.p1 = (float)cam.p1; c.p2 = (float)cam.p2;
        c.p1x2 = 2.0f * c.p1; c.p2x2 = 2.0f * c.p2;
        c.t0 = (float)cam.t[0]; c.t1 = (float)cam.t[1]; c.t2 = (float)cam.t[2];
        for (int l = 0; l < FM3D_MAX_LEVELS; l++) {
            const double scale = 1.0 / (double)(1 << l);        // actual_scale_ (normaloptimizer.cpp:226-241)
            c.sfx[l] = (float)(scale * cam.fx); c.sfy[l] = (float)(scale * cam.fy);
            c.scx[l] = (float)(scale * cam.cx); c.scy[l] = (float)(scale * cam.cy);
        }
    }
    int nt = ctx->opt_normals_threads;
    nt = nt < 128 ? 128 : (nt > FAST_NT ? FAST_NT : (nt & ~63));
    // the opt-in maximum covers static + dynamic shared memory: the control structs are static
    const size_t smem_max = ctx->prop.sharedMemPerBlockOptin - fast_static_bytes();
    const size_t tail = 128;

    // window per level: the warp of the disc is close to a similarity, 1.3x the scaled radius
    // plus margins
    auto plan_windows = [&](size_t budget) -> size_t {
        size_t win_bytes = 0;
        for (int l = 0; l <= A.pyr.levels; l++) {
            const double rs = (double)A.r / (double)(1 << l);
            int half = (int)ceil(1.3 * rs) + 8;
            int ww = (2 * half + 16 + 15) & ~15, wh = 2 * half;   // +16: slack for the 16-pixel origin alignment
            while ((size_t)ww * wh > budget && ww > 32 && wh > 16) { ww = (ww * 7 / 8) & ~15; wh = wh * 7 / 8; }
            A.win_w[l] = ww; A.win_h[l] = wh;
            A.win_tma[l] = (ww <= 256 && wh <= 256) ? 1 : 0;
            if ((size_t)ww * wh > win_bytes) win_bytes = (size_t)ww * wh;
        }
        return (win_bytes + 127) & ~(size_t)127;
    };
    // Layouts, in order of preference.  Two groups per CTA (the passes of one feature hide the
    // serial LM step of the other) need the rays in global memory (L2-resident per-group scratch);
    // one group keeps everything in shared memory when it fits.
    struct Layout { int groups; bool rays_smem, i1_smem; };
    const Layout two_a = {2, false, true}, two_b = {2, false, false}, one_a = {1, true, true}, one_b = {1, false, false};
    const Layout four = {4, false, false};
    Layout order[7];
    int n_order = 0;
    const int sms = ctx->prop.multiProcessorCount;
    if (ctx->opt_normals_groups == 3 && nt % 96 == 0) { const Layout three = {3, false, false}; order[n_order++] = three; }
    if ((ctx->opt_normals_groups == 4 || (ctx->opt_normals_groups == 0 && A.n >= 5 * sms)) && nt % 128 == 0) order[n_order++] = four;
    const int want = ctx->opt_normals_groups;   // 0: automatic
    const bool prefer_two = want == 2 || (want == 0 && A.n > ctx->prop.multiProcessorCount);
    if (prefer_two) { order[n_order++] = two_a; order[n_order++] = two_b; }
    order[n_order++] = one_a;
    if (!prefer_two && want != 1) { order[n_order++] = two_a; order[n_order++] = two_b; }
    order[n_order++] = one_b;
    Layout lay = one_b;
    size_t win_bytes = 0, group_smem = 0;
    bool found = false;
    for (int k = 0; k < n_order && !found; k++) {
        const Layout c = order[k];
        const bool last = (k == n_order - 1);
        // every layout but the last resort must hold the full-size window (a shrunken window sends
        // the level-0 passes to the global-memory tap path)
        win_bytes = plan_windows(last ? smem_max - tail - 1024 : (size_t)1 << 30);
        group_smem = (win_bytes + (c.rays_smem ? sizeof(float2) * (size_t)A.mcap : 0) +
                      (c.i1_smem ? sizeof(float) * (size_t)A.mcap : 0) + tail + 127) & ~(size_t)127;
        if (group_smem * c.groups <= smem_max) { lay = c; found = true; }
    }
    if (!found) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "normals kernel does not fit in shared memory");
    A.win_bytes = (int)win_bytes;
    A.groups = lay.groups;
    A.group_smem = (int)group_smem;
    for (int l = 0; l <= A.pyr.levels; l++) {
        if (A.use_tma && A.win_tma[l]) {
            const fm3d_level& lv = A.pyr.lv[l];
            if (int rc = fm3d_encode_tmap_2d_u8(ctx, &A.tmap[l], A.pyr.base[1] + lv.off, lv.w, lv.h, lv.pitch,
                                                A.win_w[l], A.win_h[l]))
                return rc;
            if (int rc = fm3d_encode_tmap_2d_u8(ctx, &A.tmap1[l], A.pyr.base[0] + lv.off, lv.w, lv.h, lv.pitch,
                                                A.win_w[l], A.win_h[l]))
                return rc;
        }
    }
    const size_t smem = group_smem * lay.groups;
    const bool two_slot = ctx->opt_normals_pingpong && lay.groups == 4 && !lay.rays_smem && !lay.i1_smem && A.mode == 0 && nt == FAST_NT;

    auto launch = [&](auto kernel) -> int {
        FM3D_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        FM3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, nt, smem));
        if (occ < 1) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "normals kernel does not fit (smem %zu)", smem);
        int grid = ctx->prop.multiProcessorCount * occ;
        const int need = (A.n + lay.groups - 1) / lay.groups;
        if (grid > need) grid = need;
        int* ctrl = nullptr;  // [0] work counter, [1] error flag, [16..] counters
        if (int rc = fm3d_scratch(ctx, 1, 256, (void**)&ctrl)) return rc;
        ctx->n_copy++;
        FM3D_CUDA(ctx, cudaMemsetAsync(ctrl, 0, 256, ctx->stream));
        A.work_counter = ctrl;
        A.error_flag = ctrl + 1;
        A.stats = reinterpret_cast<unsigned long long*>(ctrl + 16);
        A.rays_g = nullptr; A.i1_g = nullptr;
        if (!lay.rays_smem || !lay.i1_smem) {
            const size_t slots = (size_t)grid * lay.groups;
            char* gm = nullptr;
            if (int rc = fm3d_scratch(ctx, 2, (sizeof(float2) + sizeof(float)) * (size_t)A.mcap * slots, (void**)&gm)) return rc;
            A.rays_g = (float2*)gm;
            A.i1_g = (float*)(gm + sizeof(float2) * (size_t)A.mcap * slots);
        }
        kernel<<<grid, nt, smem, ctx->stream>>>(A);
        FM3D_LAUNCH_CHECK(ctx);
        return FM3D_OK;
    };
    // four windows, rays and image-1 samples in the scratch, mode 0: two features per eight warps taking turns
    if (two_slot) {
        if (A.cost_mode == FM3D_COST_NCC) return launch(normals_pp_kernel<true>);
        return launch(normals_pp_kernel<false>);
    }
    if (A.cost_mode == FM3D_COST_NCC) {
        if (lay.rays_smem) return launch(normals_fast_kernel<true, true, true>);
        if (lay.i1_smem) return launch(normals_fast_kernel<false, true, true>);
        return launch(normals_fast_kernel<false, false, true>);
    }
    if (lay.rays_smem) return launch(normals_fast_kernel<true, true, false>);
    if (lay.i1_smem) return launch(normals_fast_kernel<false, true, false>);
    return launch(normals_fast_kernel<false, false, false>);
}

}  // namespace fm3d_normals
