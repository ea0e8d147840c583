// fm3d_normals_common.cuh -- pieces shared by the two implementations of the plane-normal search:
// fm3d_normals.cu      "faithful": fp64 geometry, forward-difference Jacobian, one pass per
//                      lmfit evaluation group (the reference's arithmetic, evaluation by evaluation)
// fm3d_normals_fast.cu "fast": fp32 offset geometry around the centre ray, analytic Jacobian,
//                      trial and Jacobian evaluated in one pass
#ifndef FM3D_NORMALS_COMMON_CUH_
#define FM3D_NORMALS_COMMON_CUH_

#include <math.h>

#include "fm3d_internal.cuh"
#include "fm3d_lm2.h"

namespace fm3d_normals {

constexpr int NT_MAX = 512;
constexpr int WIN_MAX_W = 192, WIN_MAX_H = 192;
constexpr int WIN_BYTES = WIN_MAX_W * WIN_MAX_H;
constexpr int MAX_RAY = 255;
constexpr int MAX_ROWS = 2 * MAX_RAY + 1;

enum { FLAG_NAN = 1, FLAG_BBOX = 2, FLAG_PIX = 4 };

// fast kernel: the camera constants of the pixel loops as floats (filled by run_normals_fast)
struct FastConsts {
    float k1, k2, k3, p1, p2;
    float p1x2, p2x2;                                  // 2 p1, 2 p2
    float t0, t1, t2;                                  // translation of camera 2 (the epipole, for the Jacobian)
    float sfx[FM3D_MAX_LEVELS], sfy[FM3D_MAX_LEVELS];  // scale * K per pyramid level
    float scx[FM3D_MAX_LEVELS], scy[FM3D_MAX_LEVELS];
};

struct NormalsArgs {
    fm3d_cam cam;
    FastConsts fc;
    fm3d_pyramid_desc pyr;
    const double* xyz;
    int n;
    int r;
    double eps_lmmin;
    int penalty_mode;
    int patience;
    int mode;               // 0 optimise, 1 evaluate the cost at phi_theta / level only, 2 sweep a grid of candidates
    int sweep_nphi, sweep_ntheta;       // mode 2: grid size, centred on phi_theta[f] (or the viewing ray)
    double sweep_dphi, sweep_dtheta;    // mode 2: grid steps (rad)
    int32_t* best_idx;                  // mode 2: argmin over the grid (nullable)
    double* best_cost;                  // mode 2: its cost (nullable)
    int eval_level;
    const double* phi_theta;
    double* normals;
    int32_t* status;
    int32_t* nfev;
    int32_t* npenalty;
    double* cost;
    int32_t* m_out;
    int* work_counter;
    int* error_flag;        // set to 1 if a TMA wait timed out (the kernel then falls back)
    unsigned long long* stats;  // 8 counters of executed work (fm3d_get_normals_stats), nullable
    int mcap;
    int win_w[FM3D_MAX_LEVELS], win_h[FM3D_MAX_LEVELS];
    int use_tma;
    int win_bytes;                      // fast kernel: bytes of the window buffer
    int win_tma[FM3D_MAX_LEVELS];       // fast kernel: level window loaded by one TMA tile
    int fuse_trials;                    // fast kernel: evaluate the Jacobian with the first trial
    int memo_trials;                    // fast kernel: answer coefficient-identical trials without a pass
    int level_sync;                     // two-slot kernel: 1 = the shared warps wait for the LM warp's start of a level at the end of the level set-up
    int sweep_batch;                    // fast kernel, mode 2: > 1 evaluates the candidate grid in batches of SWEEP_B per pass
    int cost_mode;                      // fast kernel: FM3D_COST_SSD (the reference) / FM3D_COST_NCC
    int groups;                         // fast kernel: independent feature pipelines per CTA (1 or 2)
    int group_smem;                     // fast kernel: bytes of shared memory per group
    float2* rays_g;
    float* i1_g;
    CUtensorMap tmap[FM3D_MAX_LEVELS];      // image 2, one per level
    CUtensorMap tmap1[FM3D_MAX_LEVELS];     // image 1 (fast kernel: image-1 samples are taken from a staged window too)
};

// Parameters of one pass, written by thread 0 and read by everybody.
template <typename G>
struct PassParams {
    G nx[3], ny[3], nz[3], mnum[3];
    int ne;        // 1 (trial) or 3 (Jacobian)
    int cmd;       // fm3d_lm_cmd, or 0 = feature finished
};

struct FeatureShared {
    fm3d_lm2 lm;
    double w[3];       // penalty weights of the evaluations of the current pass
    double P[3];
    double normal[3];
    int feature;
    int status;
    int npenalty;
    int level;
    int m;
    int wx0, wy0;      // window origin (level pixels)
    int tma_phase;
};

// ---------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------- small helpers
__device__ __forceinline__ double penalty_weight(double phi, double theta, int mode, int& entered) {
    entered = 0;
    if (mode == FM3D_PENALTY_OFF) return 1.0;
    double at, ap;
    if (mode == FM3D_PENALTY_INT_ABS) { at = (double)abs((int)theta); ap = (double)abs((int)phi); }
    else { at = fabs(theta); ap = fabs(phi); }
    const double pi = 3.14159265358979323846;
    if (at - pi / 2 > 0 || ap - pi > 0) {
        const double wt = exp(at - pi / 2) + 1;
        const double wp = exp(ap - pi + 1) + 1;
        entered = 1;
        return wp * wt;
    }
    return 1.0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Row-major walk over the clipped disc: pixel idx lives in row `row` at x offset i.
struct RowTable {
    int start[MAX_ROWS + 1];
    short ilo[MAX_ROWS];
    short jrow[MAX_ROWS];
    int nrows;
};

inline int disc_capacity(int r) {
    int m = 0;
    for (int j = -r; j <= r; j++) m += 2 * (int)floor(sqrt((double)(r * r - j * j))) + 1;
    return m;
}

// fm3d_normals_fast.cu
int run_normals_fast(fm3d_ctx* ctx, NormalsArgs& A);

}  // namespace fm3d_normals
#endif  // FM3D_NORMALS_COMMON_CUH_
