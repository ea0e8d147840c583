// fm3d_describe.cu -- K9: the descriptor of every rectified patch, one CTA per patch.
//
// Replaces DescriptorsMatcher::extractDescriptorsFromPatches
// (DescriptorsMatcher/descriptorsmatcher.cpp:133-174) for ExtractorType SIFT (:302-314): ONE keypoint per
// patch, at (floor(S/2), floor(S/2)), size = S, angle = -1, octave = 0, handed to
// DescriptorExtractor::compute.  With a provided keypoint of octave 0 cv::SIFT builds a one-octave
// pyramid without up-sampling (firstOctave = 0), so the descriptor is read from
//     base = GaussianBlur(float(patch), sigma = sqrt(1.6^2 - 0.5^2))        (createInitialImage)
// by calcSIFTDescriptor(base, pt, ori = 360 - (-1), scl = S/2, d = 4, n = 8): hist_width = 3 scl =
// 1.5 S, so the sampling window covers the whole patch and every pixel falls between the four central
// spatial cells (rbin, cbin in (1, 2)): at most 34 of the 128 values are non-zero.  That geometry is what
// the reference's keypoint implies for any S; the host checks it instead of assuming it.
//
// Per patch, all on chip: u8 -> float, separable 13-tap (ksize = round(8 sigma + 1) | 1) symmetric
// blur with BORDER_REFLECT_101, central differences, cv::fastAtan2's polynomial, exp weight,
// OpenCV's trilinear split (v_r1 = mag*rbin, v_r0 = mag - v_r1, ...) accumulated in per-thread
// PRIVATE columns of shared memory (bank = lane: conflict-free for any orientation bin, no atomics,
// fixed summation order => deterministic), fixed-order reduction, then the descriptor
// normalisation (clip at 0.2 |h|, scale to 512, saturate to u8).
#include "fm3d_internal.cuh"

#include <math.h>

namespace {

constexpr int DESC_NT = 512;
constexpr int DESC_KHALF = 6;        // ksize 13 for sigma 1.5199: cvRound(sigma * 8 + 1) | 1
constexpr int DESC_SLOTS = 34;       // 2 x 2 central cells x 8 orientations + 2 spill slots (see o0 == -1 below)
constexpr int DESC_ROWS = 9;         // accumulator rows per thread: 8 orientation bins x float4 (cells) + the spill row
constexpr int DESC_HALO = 8;         // reflected columns kept on either side of a patch row (>= DESC_KHALF, 16-byte aligned)

struct DescArgs {
    int S, n;
    float kern[DESC_KHALF + 1];      // kern[i] = coefficient at distance i from the centre tap
    float cos_t, sin_t;              // already divided by hist_width
    float ori, bins_per_rad, exp_scale;
    int pt;                          // keypoint pixel (both coordinates)
    int bufA_floats;                 // S * (S + 2 DESC_HALO)
};

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// cv::fastAtan2 (degrees in [0, 360), max error 0.3 deg): the polynomial the reference's
// orientation bins are computed from.
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = 57.29577951308232f;
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = 2.220446049250313e-16f;
    const float ax = fabsf(x), ay = fabsf(y);
    const bool ge = ax >= ay;
    // one approximate division (2 ulp): the polynomial itself is only good to 0.3 degrees, and a bin
    // decision can only change within ~1e-5 degrees of a bin edge
    const float c = __fdividef(ge ? ay : ax, __fadd_rn(ge ? ax : ay, eps)), c2 = __fmul_rn(c, c);
    float a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    if (!ge) a = __fsub_rn(90.0f, a);
    if (x < 0) a = __fsub_rn(180.0f, a);
    if (y < 0) a = __fsub_rn(360.0f, a);
    return a;
}

// accumulator slot of descriptor element (i, j, o), or -1 if the reference's keypoint never fills it
__device__ __forceinline__ int desc_slot(int i, int j, int o) {
    if (i < 1 || i > 2) return -1;
    if (j == 0) return o == 1 ? 32 + (i - 1) : -1;
    if (j > 2) return -1;
    return ((i - 1) * 2 + (j - 1)) * 8 + o;
}
// k-th populated slot in descriptor order (the order calcSIFTDescriptor sums the squares in)
__device__ __forceinline__ int desc_order(int k) {
    // row i = 1: slot 32, then slots 0..15; row i = 2: slot 33, then slots 16..31
    if (k == 0) return 32;
    if (k <= 16) return k - 1;
    if (k == 17) return 33;
    return k - 2;
}

__device__ __forceinline__ float sqrt_approx(float a) {   // MUFU.RSQ-based, 2^-22 relative
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

// i / S and i % S for 0 <= i < S*S <= 2^15 without the integer-division sequence
__device__ __forceinline__ void divmod_small(int i, int S, float inv_S, int& q, int& rem) {
    q = __float2int_rd(((float)i + 0.5f) * inv_S);
    rem = i - q * S;
}

__global__ void __launch_bounds__(DESC_NT, 1)
describe_sift_kernel(const DescArgs A, const uint8_t* __restrict__ patches, float* __restrict__ desc) {
    extern __shared__ __align__(16) float sm[];
    const int S = A.S, S2 = S * S, tid = threadIdx.x;
    const int SA = S + 2 * DESC_HALO;            // bufA row stride: DESC_HALO reflected columns on either side
    float* bufA = sm;                            // float(patch) with column halo; later the blurred image (stride S)
    float* bufB = sm + A.bufA_floats;            // row pass with DESC_KHALF halo rows above and below; later the accumulators
    __shared__ float part[DESC_NT / 32][DESC_SLOTS];
    __shared__ float hist[DESC_SLOTS];
    const size_t f = blockIdx.x;
    const uint8_t* __restrict__ src = patches + f * (size_t)S2;
    const float inv_S = 1.0f / (float)S;

    // u8 -> float (16-byte loads when rows and the patch start are 16-byte aligned)
    if ((S & 15) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        const int qpr = S >> 4;                  // 16-pixel groups per row
        for (int q = tid; q < (S2 >> 4); q += DESC_NT) {
            const int r = q / qpr, c = (q - r * qpr) << 4;
            const uint4 v = s4[q];
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            float4* dst = reinterpret_cast<float4*>(bufA + r * SA + DESC_HALO + c);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                float4 o;
                o.x = fm3d_u8f(w[k] & 255u); o.y = fm3d_u8f((w[k] >> 8) & 255u);
                o.z = fm3d_u8f((w[k] >> 16) & 255u); o.w = fm3d_u8f(w[k] >> 24);
                dst[k] = o;
            }
        }
    } else {
        for (int i = tid; i < S2; i += DESC_NT) {
            int r, c;
            divmod_small(i, S, inv_S, r, c);
            bufA[r * SA + DESC_HALO + c] = fm3d_u8f(src[i]);
        }
    }
    // BORDER_REFLECT_101 columns: -k <- k, S-1+k <- S-1-k (read from global: no barrier needed before)
    for (int i = tid; i < 2 * DESC_KHALF * S; i += DESC_NT) {
        const int r = i / (2 * DESC_KHALF), k = i - r * (2 * DESC_KHALF);
        const int kk = k < DESC_KHALF ? k + 1 : k - DESC_KHALF + 1;            // 1..KHALF
        const int cdst = k < DESC_KHALF ? -kk : S - 1 + kk, csrc = k < DESC_KHALF ? kk : S - 1 - kk;
        bufA[r * SA + DESC_HALO + cdst] = fm3d_u8f(src[r * S + csrc]);
    }
    __syncthreads();
    // row pass (symmetric form: k0 x0 + sum_i k_i (x_-i + x_+i)) -> bufB rows KHALF .. KHALF+S-1.  Four outputs per
    // thread from five aligned 16-byte loads (columns c-8 .. c+11) instead of 13 loads per output: the kernel is
    // bound by the shared-memory pipe (profiles/r01f_describe_kernel_ncu_*).
    if ((S & 3) == 0) {
        const int S4 = S >> 2;
        for (int i = tid; i < S * S4; i += DESC_NT) {
            int r, c4;
            divmod_small(i, S4, 1.0f / (float)S4, r, c4);
            const float4* q = reinterpret_cast<const float4*>(bufA + r * SA + DESC_HALO + 4 * c4 - 8);
            float w[20];
#pragma unroll
            for (int k = 0; k < 5; k++) { const float4 v = q[k]; w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w; }
            float4 o;
            float* op = &o.x;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float s = __fmul_rn(A.kern[0], w[8 + j]);
#pragma unroll
                for (int k = 1; k <= DESC_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(w[8 + j - k], w[8 + j + k]), s);
                op[j] = s;
            }
            *reinterpret_cast<float4*>(bufB + (r + DESC_KHALF) * S + 4 * c4) = o;
        }
    } else {
        for (int i = tid; i < S2; i += DESC_NT) {
            int r, c;
            divmod_small(i, S, inv_S, r, c);
            const float* px = bufA + r * SA + DESC_HALO + c;
            float s = __fmul_rn(A.kern[0], px[0]);
#pragma unroll
            for (int k = 1; k <= DESC_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(px[-k], px[k]), s);
            bufB[i + DESC_KHALF * S] = s;
        }
    }
    __syncthreads();
    // BORDER_REFLECT_101 rows of the row-pass image
    for (int i = tid; i < 2 * DESC_KHALF * S; i += DESC_NT) {
        int k, c;
        divmod_small(i, S, inv_S, k, c);
        const int kk = k < DESC_KHALF ? k + 1 : k - DESC_KHALF + 1;
        const int rdst = k < DESC_KHALF ? -kk : S - 1 + kk, rsrc = k < DESC_KHALF ? kk : S - 1 - kk;
        bufB[(rdst + DESC_KHALF) * S + c] = bufB[(rsrc + DESC_KHALF) * S + c];
    }
    __syncthreads();
    // column pass -> bufA (stride S): four consecutive rows per thread from 16 loads (rows r-6 .. r+9 of one column)
    if ((S & 3) == 0) {
        const int R4 = S >> 2;
        for (int i = tid; i < R4 * S; i += DESC_NT) {
            int r4, c;
            divmod_small(i, S, inv_S, r4, c);
            const float* px = bufB + (4 * r4) * S + c;          // row 4 r4 - KHALF of the padded row-pass image
            float w[16];
#pragma unroll
            for (int k = 0; k < 16; k++) w[k] = px[k * S];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                float s = __fmul_rn(A.kern[0], w[6 + j]);
#pragma unroll
                for (int k = 1; k <= DESC_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(w[6 + j - k], w[6 + j + k]), s);
                bufA[(4 * r4 + j) * S + c] = s;
            }
        }
    } else {
        for (int i = tid; i < S2; i += DESC_NT) {
            const float* px = bufB + i + DESC_KHALF * S;
            float s = __fmul_rn(A.kern[0], px[0]);
#pragma unroll
            for (int k = 1; k <= DESC_KHALF; k++) s = fmaf(A.kern[k], __fadd_rn(px[-k * S], px[k * S]), s);
            bufA[i] = s;
        }
    }
    __syncthreads();
    // private accumulators: one float4 (the four central cells 2 dr + dc) per orientation bin and thread,
    // row 8 = the two spill slots
    float4* acc = reinterpret_cast<float4*>(bufB);
#pragma unroll
    for (int b = 0; b < DESC_ROWS; b++) acc[b * DESC_NT + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    // gradient histogram over the interior pixels (r, c in [1, S-2]), row-major as OpenCV walks them
    const int SI = S - 2;
    const float inv_SI = 1.0f / (float)SI;
    for (int i = tid; i < SI * SI; i += DESC_NT) {
        int rr, cc;
        divmod_small(i, SI, inv_SI, rr, cc);
        const int r = rr + 1, c = cc + 1;
        const float fi = (float)(r - A.pt), fj = (float)(c - A.pt);
        const float c_rot = __fsub_rn(__fmul_rn(fj, A.cos_t), __fmul_rn(fi, A.sin_t));
        const float r_rot = __fadd_rn(__fmul_rn(fj, A.sin_t), __fmul_rn(fi, A.cos_t));
        const float rbin = __fsub_rn(__fadd_rn(r_rot, 2.0f), 0.5f) - 1.0f;   // host-checked: floor(rbin) == 1: exact
        const float cbin = __fsub_rn(__fadd_rn(c_rot, 2.0f), 0.5f) - 1.0f;
        const float* p = bufA + r * S + c;
        const float dx = __fsub_rn(p[1], p[-1]), dy = __fsub_rn(p[-S], p[S]);
        // weight in [0.94, 1] (|rot| < 0.5): ex2.approx is good to 1e-7 here, like cv::exp's table
        const float w = __expf(__fmul_rn(__fadd_rn(__fmul_rn(c_rot, c_rot), __fmul_rn(r_rot, r_rot)), A.exp_scale));
        const float mag = __fmul_rn(sqrt_approx(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))), w);
        float obin = __fmul_rn(__fsub_rn(fast_atan2_deg(dy, dx), A.ori), A.bins_per_rad);
        const float of = floorf(obin);
        obin = __fsub_rn(obin, of);
        // calcSIFTDescriptor wraps o0 ONCE (`if (o0 < 0) o0 += n`).  The reference's keypoint has angle -1,
        // i.e. ori = 361: orientations below 1 degree give floor(obin) = -9 and leave o0 = -1.  OpenCV
        // addresses its (d+2)(d+2)(n+2) histogram flat, so hist[idx - 1] is slot n+1 of the previous
        // COLUMN cell, which the circular fold adds to that cell's orientation bin 1; the v1 part
        // (slot o0 + 1 = 0) stays in the proper cell.  Reproduced: `spill`.
        int o0 = (int)of;
        o0 = o0 < 0 ? o0 + 8 : o0;
        o0 = o0 >= 8 ? o0 - 8 : o0;
        const bool spill = o0 < 0;
        const float v_r1 = __fmul_rn(mag, rbin), v_r0 = __fsub_rn(mag, v_r1);
        const float v_rc11 = __fmul_rn(v_r1, cbin), v_rc10 = __fsub_rn(v_r1, v_rc11);
        const float v_rc01 = __fmul_rn(v_r0, cbin), v_rc00 = __fsub_rn(v_r0, v_rc01);
        float4 v1, v0;                            // .x .y .z .w = cells (0,0) (0,1) (1,0) (1,1)
        v1.x = __fmul_rn(v_rc00, obin); v0.x = __fsub_rn(v_rc00, v1.x);
        v1.y = __fmul_rn(v_rc01, obin); v0.y = __fsub_rn(v_rc01, v1.y);
        v1.z = __fmul_rn(v_rc10, obin); v0.z = __fsub_rn(v_rc10, v1.z);
        v1.w = __fmul_rn(v_rc11, obin); v0.w = __fsub_rn(v_rc11, v1.w);
        if (!spill) {
            o0 &= 7;                              // memory safety only
            float4* a0 = acc + o0 * DESC_NT + tid;
            float4* a1 = acc + ((o0 + 1) & 7) * DESC_NT + tid;      // slot n folds onto bin 0
            float4 t = *a0;
            t.x = __fadd_rn(t.x, v0.x); t.y = __fadd_rn(t.y, v0.y); t.z = __fadd_rn(t.z, v0.z); t.w = __fadd_rn(t.w, v0.w);
            *a0 = t;
            t = *a1;
            t.x = __fadd_rn(t.x, v1.x); t.y = __fadd_rn(t.y, v1.y); t.z = __fadd_rn(t.z, v1.z); t.w = __fadd_rn(t.w, v1.w);
            *a1 = t;
        } else {
            // v1 -> bin 0 of the proper cells; v0 -> bin 1 one column to the left: cell (dr,1) -> cell (dr,0),
            // cell (dr,0) -> spill slot dr
            float4* a = acc + tid;
            float4 t = *a;
            t.x = __fadd_rn(t.x, v1.x); t.y = __fadd_rn(t.y, v1.y); t.z = __fadd_rn(t.z, v1.z); t.w = __fadd_rn(t.w, v1.w);
            *a = t;
            a = acc + DESC_NT + tid;
            t = *a;
            t.x = __fadd_rn(t.x, v0.y); t.z = __fadd_rn(t.z, v0.w);
            *a = t;
            a = acc + 8 * DESC_NT + tid;
            t = *a;
            t.x = __fadd_rn(t.x, v0.x); t.y = __fadd_rn(t.y, v0.z);
            *a = t;
        }
    }
    // fixed-order reduction: lanes (xor tree), then warps in order
    const int lane = tid & 31, wid = tid >> 5;
#pragma unroll 3
    for (int b = 0; b < DESC_ROWS; b++) {
        float4 s = acc[b * DESC_NT + tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s.x += __shfl_xor_sync(0xffffffffu, s.x, o); s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
            s.z += __shfl_xor_sync(0xffffffffu, s.z, o); s.w += __shfl_xor_sync(0xffffffffu, s.w, o);
        }
        if (lane == 0) {
            if (b < 8) { part[wid][b] = s.x; part[wid][8 + b] = s.y; part[wid][16 + b] = s.z; part[wid][24 + b] = s.w; }
            else { part[wid][32] = s.x; part[wid][33] = s.y; }
        }
    }
    __syncthreads();
    if (tid < DESC_SLOTS) {
        float s = 0.0f;
        for (int w = 0; w < DESC_NT / 32; w++) s += part[w][tid];
        hist[tid] = s;
    }
    __syncthreads();
    // normalisation (calcSIFTDescriptor's tail): sequential float sums over the 128 values in order
    __shared__ float scale_thr[2];
    if (tid == 0) {
        // dst order: row 1 = {spill slot 32 at (1,0,1)}, cells 0, 1; row 2 = {slot 33}, cells 2, 3 (zeros add nothing)
        float nrm2 = 0.0f;
        for (int k = 0; k < DESC_SLOTS; k++) { const float h = hist[desc_order(k)]; nrm2 = __fadd_rn(nrm2, __fmul_rn(h, h)); }
        const float thr = __fmul_rn(__fsqrt_rn(nrm2), 0.2f);
        nrm2 = 0.0f;
        for (int k = 0; k < DESC_SLOTS; k++) {
            const float val = fminf(hist[desc_order(k)], thr);
            nrm2 = __fadd_rn(nrm2, __fmul_rn(val, val));
        }
        scale_thr[0] = __fdiv_rn(512.0f, fmaxf(__fsqrt_rn(nrm2), 1.1920929e-07f));
        scale_thr[1] = thr;
    }
    __syncthreads();
    if (tid < 128) {
        // dst[(i*4 + j)*8 + o] <- hist cell (i, j); the populated cells are i, j in {1, 2} (+ spill slots)
        const int i = tid >> 5, j = (tid >> 3) & 3, o = tid & 7;
        const int b = desc_slot(i, j, o);
        float out = 0.0f;
        if (b >= 0) out = fminf(fmaxf(rintf(__fmul_rn(fminf(hist[b], scale_thr[1]), scale_thr[0])), 0.0f), 255.0f);   // saturate_cast<uchar>
        desc[f * 128 + tid] = out;
    }
}

// The keypoint of extractDescriptorsFromPatches for patch edge S -> kernel arguments; false if
// the sampling window would not map every interior pixel between the four central cells.
bool desc_args(int S, int n, DescArgs& A) {
    A.S = S; A.n = n;
    const double sigma = sqrt(fmax(1.6 * 1.6 - 0.5 * 0.5, 0.01));       // createInitialImage, no up-sampling
    double t[2 * DESC_KHALF + 1], sum = 0;
    for (int i = 0; i <= 2 * DESC_KHALF; i++) {                          // cv::getGaussianKernel(13, sigma, CV_32F)
        const double x = i - DESC_KHALF;
        t[i] = exp(-0.5 / (sigma * sigma) * x * x);
        sum += t[i];
    }
    for (int i = 0; i <= DESC_KHALF; i++) A.kern[i] = (float)(t[DESC_KHALF + i] / sum);
    const int ksize = ((int)lrint(sigma * 8 + 1)) | 1;
    if (ksize != 2 * DESC_KHALF + 1) return false;
    A.pt = (int)lrintf((float)(S / 2));
    const float ori = 360.0f - (-1.0f);                                  // angle = 360 - kpt.angle
    const float scl = (float)S * 0.5f;                                   // kpt.size * scale * 0.5
    float cos_t = cosf(ori * (float)(M_PI / 180)), sin_t = sinf(ori * (float)(M_PI / 180));
    const float hist_width = 3.0f * scl;
    A.ori = ori;
    A.bins_per_rad = 8 / 360.0f;
    A.exp_scale = -1.0f / (4 * 4 * 0.5f);
    A.cos_t = cos_t / hist_width;
    A.sin_t = sin_t / hist_width;
    int radius = (int)lrintf(hist_width * 1.4142135623730951f * 5 * 0.5f);
    const int diag = (int)sqrt((double)S * S + (double)S * S);
    if (radius > diag) radius = diag;
    // every interior pixel must be inside the window and between the central cells
    const int lo = 1 - A.pt, hi = S - 2 - A.pt;
    if (-radius > lo || radius < hi) return false;
    const int ends[2] = {lo, hi};
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            const float i = (float)ends[a], j = (float)ends[b];
            const float c_rot = j * A.cos_t - i * A.sin_t, r_rot = j * A.sin_t + i * A.cos_t;
            const float rbin = r_rot + 2 - 0.5f, cbin = c_rot + 2 - 0.5f;
            if (!(rbin >= 1.0f && rbin < 2.0f && cbin >= 1.0f && cbin < 2.0f)) return false;
        }
    return true;
}

}  // namespace

extern "C" {

int fm3d_describe_patches_sift_dev(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, float* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && S >= 8 && (n == 0 || (patches && descriptors)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    DescArgs A;
    if (!desc_args(S, n, A))
        return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "patch edge %d: the SIFT window of the reference's keypoint does not cover the patch", S);
    A.bufA_floats = (S * (S + 2 * DESC_HALO) + 3) & ~3;      // the float4 accumulators follow: keep 16-byte alignment
    const size_t rowpass = (size_t)(S + 2 * DESC_KHALF) * S * sizeof(float);
    const size_t accb = (size_t)DESC_ROWS * DESC_NT * sizeof(float4);
    const size_t smem = (size_t)A.bufA_floats * sizeof(float) + (rowpass > accb ? rowpass : accb);
    if (smem + 4096 > ctx->prop.sharedMemPerBlockOptin)
        return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "patch edge %d needs %zu bytes of shared memory", S, smem);
    FM3D_CUDA(ctx, cudaFuncSetAttribute(describe_sift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    describe_sift_kernel<<<n, DESC_NT, smem, ctx->stream>>>(A, patches, descriptors);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_describe_patches_sift(fm3d_ctx* ctx, const uint8_t* patches, int n, int S, float* descriptors) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, n >= 0 && S >= 8 && (n == 0 || (patches && descriptors)));
    if (n == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bp = (size_t)n * S * S, bd = sizeof(float) * 128 * (size_t)n;
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, al(bp) + al(bd), (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d, patches, bp)) return rc;
    if (int rc = fm3d_describe_patches_sift_dev(ctx, (const uint8_t*)d, n, S, (float*)(d + al(bp)))) return rc;
    if (int rc = fm3d_d2h(ctx, descriptors, d + al(bp), bd)) return rc;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FM3D_OK;
}

}  // extern "C"
