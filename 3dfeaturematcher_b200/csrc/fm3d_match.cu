// fm3d_match.cu -- K1/K2: exact brute-force 2-nearest-neighbour descriptor matching with a
// fused top-2 epilogue, Lowe ratio test and mutual-best flag.
//
// Replaces matcher_->knnMatch(a, b, matches, 2) + the NNDR loop of
// DescriptorsMatcher::compareWithNNDR / ::compare / ::crosscompare
// (DescriptorsMatcher/descriptorsmatcher.cpp:74-131).  The reference searches with FLANN
// (approximate, randomised); the north star asks for the exact search with cv::BFMatcher
// semantics: L2 distances square-rooted in float, Hamming counts as float, ties broken by the
// lower train index, ratio test in double on the float distances.
//
// Three kernels produce per-(query, train-split) partial top-2 lists that one finalise kernel
// merges lexicographically by (distance, index):
//
//  match_tc_kernel   float descriptors that are integer-valued in [0,255] with dim 128 (SIFT):
//                    distance contraction on the 5th-gen tensor cores.  The operands are
//                    re-tiled once into the UMMA K-major core-matrix layout as bf16 with K
//                    extended from 128 to 144: A = [-2q | qq2 qq1 qq0 65536 256 1 0..],
//                    B = [t | 65536 256 1 tt2 tt1 tt0 0..] (|q|^2, |t|^2 split in base 256), so
//                    the fp32 accumulator in TMEM holds |q|^2 + |t|^2 - 2 q.t = d^2 exactly (all
//                    partial sums are integers below 2^24).  Warp-specialised: one thread
//                    streams 72 KB train tiles with cp.async.bulk (UBLKCP) into a 2-stage smem
//                    ring, one thread issues tcgen05.mma (M128 N256 K16 x 9) into a double-
//                    buffered 2 x 256-column TMEM accumulator, four warps drain it with
//                    tcgen05.ld (one query row per thread) keeping a running top-2 in registers.
//  match_sp_kernel   real-valued float descriptors (SURF, RootSIFT, ...; dim <= 128, dim % 4 == 0, at least
//                    2^22 pairs): bf16 hi/lo split contraction on the tensor cores as a FILTER that keeps the
//                    four smallest approximate distances per query and train split, then sp_refine_kernel
//                    decides exactly with match_f32_kernel's arithmetic and proves that no other train
//                    descriptor can win; queries it cannot prove go to match_f32_kernel.  Same results as
//                    the exact path, bit for bit.
//  match_f32_kernel  any other float descriptors: exact CUDA-core path, sum of squared
//                    differences accumulated in ascending dimension order with separately
//                    rounded multiply and add (bit-identical to the scalar oracle).
//  match_ham_kernel  binary descriptors: xor + popc on 32-bit words, train tile staged in smem
//                    with 128-bit loads and read back with 128-bit broadcast loads; packed
//                    (distance << 22 | index) keys make the top-2 update three integer min/max.
#include <cuda_bf16.h>
#include <math.h>

#include "fm3d_internal.cuh"
#include "fm3d_match_pieces.h"

namespace {

struct Cand {
    float d;   // squared L2 / Hamming count
    int idx;
};

__device__ __forceinline__ bool cand_less(float d, int i, float d2, int i2) {
    return d < d2 || (d == d2 && i < i2);
}

__device__ __forceinline__ void top2_insert(float d, int i, float& b0, int& i0, float& b1, int& i1) {
    if (cand_less(d, i, b0, i0)) { b1 = b0; i1 = i0; b0 = d; i0 = i; }
    else if (cand_less(d, i, b1, i1)) { b1 = d; i1 = i; }
}

// ------------------------------------------------------------------------------------------
// Generic exact fp32 path.  CTA tile: 64 queries x 64 train, 256 threads, 4x4 per thread.
// ------------------------------------------------------------------------------------------
constexpr int FT = 64, FK = 16;

__global__ void __launch_bounds__(256)
match_f32_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, int dim,
                 int tiles_per_split, Cand* __restrict__ partial) {
    __shared__ float sq[FK][FT + 1];
    __shared__ float st[FK][FT + 1];
    __shared__ float s_d[FT][16][2];
    __shared__ int s_i[FT][16][2];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;  // tx: train quad, ty: query quad
    const int q0 = blockIdx.x * FT;
    const int split = blockIdx.y;
    const int nt_tiles = (nt + FT - 1) / FT;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(nt_tiles, tile_lo + tiles_per_split);
    float b0[4], b1[4];
    int i0[4], i1[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { b0[a] = INFINITY; b1[a] = INFINITY; i0[a] = 0x7fffffff; i1[a] = 0x7fffffff; }
    for (int tile = tile_lo; tile < tile_hi; tile++) {
        const int t0 = tile * FT;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
        for (int k0 = 0; k0 < dim; k0 += FK) {
            __syncthreads();
            for (int i = tid; i < FT * FK; i += 256) {
                const int r = i / FK, k = i - r * FK;
                sq[k][r] = (q0 + r < nq && k0 + k < dim) ? q[(size_t)(q0 + r) * dim + k0 + k] : 0.f;
                st[k][r] = (t0 + r < nt && k0 + k < dim) ? t[(size_t)(t0 + r) * dim + k0 + k] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < FK; k++) {
                float qa[4], tb[4];
#pragma unroll
                for (int a = 0; a < 4; a++) { qa[a] = sq[k][ty * 4 + a]; tb[a] = st[k][tx * 4 + a]; }
#pragma unroll
                for (int a = 0; a < 4; a++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        const float d = __fsub_rn(qa[a], tb[b]);
                        acc[a][b] = __fadd_rn(acc[a][b], __fmul_rn(d, d));
                    }
            }
        }
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) {
                const int j = t0 + tx * 4 + b;
                if (j < nt) top2_insert(acc[a][b], j, b0[a], i0[a], b1[a], i1[a]);
            }
    }
    // merge the 16 column-owners of every query row
#pragma unroll
    for (int a = 0; a < 4; a++) {
        s_d[ty * 4 + a][tx][0] = b0[a]; s_d[ty * 4 + a][tx][1] = b1[a];
        s_i[ty * 4 + a][tx][0] = i0[a]; s_i[ty * 4 + a][tx][1] = i1[a];
    }
    __syncthreads();
    if (tid < FT && q0 + tid < nq) {
        float m0 = INFINITY, m1 = INFINITY;
        int j0 = 0x7fffffff, j1 = 0x7fffffff;
        for (int c = 0; c < 16; c++) {
            top2_insert(s_d[tid][c][0], s_i[tid][c][0], m0, j0, m1, j1);
            top2_insert(s_d[tid][c][1], s_i[tid][c][1], m0, j0, m1, j1);
        }
        Cand* o = partial + ((size_t)split * nq + q0 + tid) * 2;
        o[0].d = m0; o[0].idx = j0;
        o[1].d = m1; o[1].idx = j1;
    }
}

// ------------------------------------------------------------------------------------------
// Hamming path.  128 threads, 2 queries per thread, train tile of 128 descriptors in smem.
// ------------------------------------------------------------------------------------------
template <int WORDS>  // 32-bit words per descriptor (8 = ORB, 16 = BRISK/FREAK)
__global__ void __launch_bounds__(128)
match_ham_kernel(const uint32_t* __restrict__ q, int nq, const uint32_t* __restrict__ t, int nt,
                 int tiles_per_split, uint32_t* __restrict__ partial) {
    constexpr int TT = 128;
    __shared__ __align__(16) uint32_t s_t[TT * WORDS];
    const int tid = threadIdx.x;
    const int qa = blockIdx.x * 256 + tid, qb = qa + 128;
    const int split = blockIdx.y;
    const int nt_tiles = (nt + TT - 1) / TT;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(nt_tiles, tile_lo + tiles_per_split);
    uint32_t A[WORDS], B[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; w += 4) {
        const uint4 va = qa < nq ? *reinterpret_cast<const uint4*>(q + (size_t)qa * WORDS + w) : make_uint4(0, 0, 0, 0);
        const uint4 vb = qb < nq ? *reinterpret_cast<const uint4*>(q + (size_t)qb * WORDS + w) : make_uint4(0, 0, 0, 0);
        A[w] = va.x; A[w + 1] = va.y; A[w + 2] = va.z; A[w + 3] = va.w;
        B[w] = vb.x; B[w + 1] = vb.y; B[w + 2] = vb.z; B[w + 3] = vb.w;
    }
    uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu, c0 = 0xffffffffu, c1 = 0xffffffffu;
    for (int tile = tile_lo; tile < tile_hi; tile++) {
        const int t0 = tile * TT;
        __syncthreads();
        for (int i = tid; i < TT * WORDS / 4; i += 128) {
            const int r = (i * 4) / WORDS;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (t0 + r < nt) v = *reinterpret_cast<const uint4*>(t + (size_t)t0 * WORDS + (size_t)i * 4);
            *reinterpret_cast<uint4*>(&s_t[i * 4]) = v;
        }
        __syncthreads();
        const int cnt = min(TT, nt - t0);
#pragma unroll 4
        for (int j = 0; j < cnt; j++) {
            int da = 0, db = 0;
#pragma unroll
            for (int w = 0; w < WORDS; w += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(&s_t[j * WORDS + w]);  // broadcast
                da += __popc(A[w] ^ v.x) + __popc(A[w + 1] ^ v.y) + __popc(A[w + 2] ^ v.z) + __popc(A[w + 3] ^ v.w);
                db += __popc(B[w] ^ v.x) + __popc(B[w + 1] ^ v.y) + __popc(B[w + 2] ^ v.z) + __popc(B[w + 3] ^ v.w);
            }
            const uint32_t ka = ((uint32_t)da << 22) | (uint32_t)(t0 + j);
            const uint32_t kb = ((uint32_t)db << 22) | (uint32_t)(t0 + j);
            uint32_t m = min(ka, a0); a1 = min(a1, max(ka, a0)); a0 = m;
            m = min(kb, c0); c1 = min(c1, max(kb, c0)); c0 = m;
        }
    }
    if (qa < nq) { uint32_t* o = partial + ((size_t)split * nq + qa) * 2; o[0] = a0; o[1] = a1; }
    if (qb < nq) { uint32_t* o = partial + ((size_t)split * nq + qb) * 2; o[0] = c0; o[1] = c1; }
}

__global__ void finalize_f32_kernel(const Cand* __restrict__ partial, int nsplit, int nq, int nt,
                                    int32_t* __restrict__ idx, float* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float m0 = INFINITY, m1 = INFINITY;
    int j0 = 0x7fffffff, j1 = 0x7fffffff;
    for (int s = 0; s < nsplit; s++) {
        const Cand* p = partial + ((size_t)s * nq + i) * 2;
        if (p[0].idx >= 0 && p[0].idx < nt) top2_insert(p[0].d, p[0].idx, m0, j0, m1, j1);
        if (p[1].idx >= 0 && p[1].idx < nt) top2_insert(p[1].d, p[1].idx, m0, j0, m1, j1);
    }
    const bool v0 = j0 < nt, v1 = j1 < nt;
    idx[2 * i] = v0 ? j0 : -1;
    idx[2 * i + 1] = v1 ? j1 : -1;
    dist[2 * i] = v0 ? sqrtf(m0) : INFINITY;       // DMatch::distance for NORM_L2
    dist[2 * i + 1] = v1 ? sqrtf(m1) : INFINITY;
}

__global__ void finalize_ham_kernel(const uint32_t* __restrict__ partial, int nsplit, int nq, int nt,
                                    int32_t* __restrict__ idx, float* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    uint32_t a0 = 0xffffffffu, a1 = 0xffffffffu;
    for (int s = 0; s < nsplit; s++) {
        const uint32_t* p = partial + ((size_t)s * nq + i) * 2;
        for (int k = 0; k < 2; k++) {
            const uint32_t key = p[k];
            const uint32_t m = min(key, a0); a1 = min(a1, max(key, a0)); a0 = m;
        }
    }
    const bool v0 = a0 != 0xffffffffu, v1 = a1 != 0xffffffffu;
    idx[2 * i] = v0 ? (int)(a0 & 0x3fffffu) : -1;
    idx[2 * i + 1] = v1 ? (int)(a1 & 0x3fffffu) : -1;
    dist[2 * i] = v0 ? (float)(a0 >> 22) : INFINITY;
    dist[2 * i + 1] = v1 ? (float)(a1 >> 22) : INFINITY;
}

// NNDR filter (descriptorsmatcher.cpp:119-129) with order-preserving compaction by one CTA.
__global__ void __launch_bounds__(1024)
nndr_kernel(const int32_t* __restrict__ idx, const float* __restrict__ dist, int nq, double eps,
            int32_t* __restrict__ qidx, int32_t* __restrict__ tidx, float* __restrict__ dout,
            int* __restrict__ nmatch) {
    __shared__ int warp_cnt[32];
    __shared__ int base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (int start = 0; start < nq; start += blockDim.x) {
        const int i = start + tid;
        bool keep = false;
        if (i < nq && idx[2 * i] >= 0 && idx[2 * i + 1] >= 0)
            keep = (double)dist[2 * i] <= eps * (double)dist[2 * i + 1];
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < wid; w++) off += warp_cnt[w];
        off += __popc(bal & ((1u << lane) - 1u));
        if (keep) { qidx[off] = i; tidx[off] = idx[2 * i]; dout[off] = dist[2 * i]; }
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); w++) tot += warp_cnt[w];
            base_s += tot;
        }
        __syncthreads();
    }
    if (tid == 0) *nmatch = base_s;
}

__global__ void mutual_kernel(const int32_t* __restrict__ qidx, const int32_t* __restrict__ tidx,
                              const int* __restrict__ nmatch, const int32_t* __restrict__ idx_ba,
                              uint8_t* __restrict__ mutual) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= *nmatch) return;
    mutual[i] = idx_ba[2 * tidx[i]] == qidx[i] ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// Tensor-core path (tcgen05 + TMEM + bulk-copy pipeline)
// ------------------------------------------------------------------------------------------
constexpr int TC_DIM = 128;
constexpr int TC_KCHUNKS = 18;                      // 16-byte K chunks per row: 16 data + 2 extras
constexpr int TC_GROUP_BYTES = TC_KCHUNKS * 128;    // 8 rows x 18 chunks x 16 B = 2304
constexpr int TC_M = 128, TC_N = 256;
constexpr int TC_A_BYTES = (TC_M / 8) * TC_GROUP_BYTES;   // 36864
constexpr int TC_B_BYTES = (TC_N / 8) * TC_GROUP_BYTES;   // 73728
constexpr int TC_STAGES = 2;
constexpr int TC_SMEM = TC_A_BYTES + TC_STAGES * TC_B_BYTES + 256;
#ifndef FM3D_TC_EPI_WARPS
#define FM3D_TC_EPI_WARPS 8                         // epilogue warps: 4 (one per TMEM lane quarter) or 8 (two per quarter, half the columns each)
#endif
constexpr int TC_EPI_WARPS = FM3D_TC_EPI_WARPS;
constexpr int TC_EPI_HALVES = TC_EPI_WARPS / 4;     // column halves of a tile, one per epilogue warp of a lane quarter
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

// Re-tile n x 128 float descriptors into the UMMA K-major no-swizzle core-matrix layout
// (8 rows x 16 B core matrices; 18 K-chunks of a row group contiguous) as bf16, rows padded
// to `n_pad`.  is_query: values are -2*q and the extras carry |q|^2; else values are t and the
// extras carry |t|^2.  not_integer is set if any value is not an integer in [0,255].
__global__ void __launch_bounds__(256)
tc_prep_kernel(const float* __restrict__ src, int n, int n_pad, int is_query,
               uint8_t* __restrict__ dst, int* __restrict__ not_integer) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);   // one warp per row
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    bool bad = false;
    if (row < n) {
        const float4 x = *reinterpret_cast<const float4*>(src + (size_t)row * TC_DIM + lane * 4);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
#pragma unroll
        for (int k = 0; k < 4; k++) bad |= !(v[k] >= 0.f && v[k] <= 255.f && v[k] == floorf(v[k]));
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(not_integer, 1);
    float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    uint8_t* g = dst + (size_t)(row >> 3) * TC_GROUP_BYTES + (row & 7) * 16;
    const float sc = is_query ? -2.f : 1.f;
    // lane holds dims 4*lane..4*lane+3 -> chunk lane/2, half lane&1
    __nv_bfloat162 p0 = __floats2bfloat162_rn(sc * v[0], sc * v[1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(sc * v[2], sc * v[3]);
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0);
    pk.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(g + (lane >> 1) * 128 + (lane & 1) * 8) = pk;
    if (lane == 0) {
        // extras: 16 bf16 = chunks 16 and 17
        const unsigned s = row < n ? (unsigned)ss : 0xffffffu;   // padded rows: |.|^2 = 2^24-1 -> never win
        const float e2 = (float)((s >> 16) & 255u), e1 = (float)((s >> 8) & 255u), e0 = (float)(s & 255u);
        float ex[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ex[k] = 0.f;
        if (is_query) { ex[0] = e2; ex[1] = e1; ex[2] = e0; ex[3] = 65536.f; ex[4] = 256.f; ex[5] = 1.f; }
        else { ex[0] = 65536.f; ex[1] = 256.f; ex[2] = 1.f; ex[3] = e2; ex[4] = e1; ex[5] = e0; }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                __nv_bfloat162 p = __floats2bfloat162_rn(ex[c * 8 + 2 * k], ex[c * 8 + 2 * k + 1]);
                w[k] = *reinterpret_cast<uint32_t*>(&p);
            }
            *reinterpret_cast<uint4*>(g + (16 + c) * 128) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (long long spin = 0; spin < (1ll << 28); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();  // a pipeline bug must fail the launch, not hang the GPU
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, no swizzle: LBO = stride between the two 16-byte K chunks of one MMA (128 B),
// SBO = stride between 8-row groups (2304 B), descriptor version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((128u >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)(((uint32_t)TC_GROUP_BYTES >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Running top-2 of one query row over 32 accumulator columns.  The group minimum is a tree (depth 5,
// FMNMX3 where the compiler finds it) rather than a 31-deep dependent chain: with one epilogue warp
// per scheduler the chain latency, not the issue rate, was what kept the tensor pipe waiting.
__device__ __forceinline__ void top2_chunk(const uint32_t (&v)[32], uint32_t taddr, int col0, float& m0, int& i0, float& m1, int& i1) {
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
        p[j] = fminf(fminf(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                     fminf(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
    const float gmin = fminf(fminf(fminf(p[0], p[1]), fminf(p[2], p[3])), fminf(fminf(p[4], p[5]), fminf(p[6], p[7])));
    if (__any_sync(0xffffffffu, gmin < m1)) {
        // columns that beat some row's second value: few; a compact warp-uniform loop re-reads them from TMEM
        // (ascending columns and strict <: the lower index wins ties) instead of a 32-way unrolled insertion
        unsigned mask = 0u;
#pragma unroll
        for (int k = 0; k < 32; k++) mask |= (__uint_as_float(v[k]) < m1) ? (1u << k) : 0u;
        unsigned um = __reduce_or_sync(0xffffffffu, mask);
        while (um) {
            const int k = __ffs(um) - 1;
            um &= um - 1u;
            uint32_t x;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(x) : "r"(taddr + (uint32_t)k));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const float d = __uint_as_float(x);
            if (d < m1) {
                const int j = col0 + k;
                if (d < m0) { m1 = m0; i1 = i0; m0 = d; i0 = j; }
                else { m1 = d; i1 = j; }
            }
        }
    }
}

// grid = (query tiles of 128, train splits).  qa / tb: re-tiled operands (tc_prep_kernel).
__global__ void __launch_bounds__(TC_THREADS, 1)
match_tc_kernel(const uint8_t* __restrict__ qa, const uint8_t* __restrict__ tb, int nq, int nt_tiles,
                int tiles_per_split, Cand* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + TC_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_A_BYTES + TC_STAGES * TC_B_BYTES);
    uint64_t* a_full = bars + 0;
    uint64_t* b_full = bars + 1;            // [TC_STAGES]
    uint64_t* b_empty = bars + 1 + TC_STAGES;
    uint64_t* acc_full = bars + 1 + 2 * TC_STAGES;   // [2]
    uint64_t* acc_empty = bars + 3 + 2 * TC_STAGES;  // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 5 + 2 * TC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, split = blockIdx.y;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(nt_tiles, tile_lo + tiles_per_split);
    const int ntiles = tile_hi - tile_lo;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < TC_STAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        const uint32_t ncols = 512;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        // ===== producer: one thread streams the query tile once and the train tiles through the ring
        if (lane == 0) {
            mbar_expect_tx(a_full, TC_A_BYTES);
            bulk_g2s(sA, qa + (size_t)qtile * TC_A_BYTES, TC_A_BYTES, a_full);
            for (int it = 0; it < ntiles; it++) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1u;
                mbar_wait(&b_empty[s], ph ^ 1u);
                mbar_expect_tx(&b_full[s], TC_B_BYTES);
                bulk_g2s(sB + (size_t)s * TC_B_BYTES, tb + (size_t)(tile_lo + it) * TC_B_BYTES, TC_B_BYTES, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread, D[tmem] (+)= A[smem] * B[smem]^T, 9 K-steps of 16
        if (lane == 0) {
            // kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N=256, M=128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            mbar_wait(a_full, 0);
            const uint32_t a_addr = s32(sA);
            for (int it = 0; it < ntiles; it++) {
                const int s = it % TC_STAGES, acc = it & 1;
                mbar_wait(&b_full[s], (uint32_t)(it / TC_STAGES) & 1u);
                mbar_wait(&acc_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = s32(sB + (size_t)s * TC_B_BYTES);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_N);
#pragma unroll
                for (int k = 0; k < TC_KCHUNKS / 2; k++)
                    umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 256), umma_smem_desc(b_addr + k * 256), idesc, k > 0);
                umma_commit(&b_empty[s]);     // smem stage reusable once these MMAs have read it
                umma_commit(&acc_full[acc]);  // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: thread <-> TMEM lane <-> query row; running top-2 in registers.  The epilogue is co-critical
        // with the MMA (latency-bound min trees and TMEM loads), so two warps share a lane quarter and take half the
        // columns of every tile each; their lists are merged by finalize like two more train splits.
        const int lane_grp = warp & 3;                  // a warp may only touch TMEM lanes 32*(warp%4)..+31
        const int half = (warp - 2) >> 2;               // 0 .. TC_EPI_HALVES-1
        const int row = lane_grp * 32 + lane;
        float m0 = INFINITY, m1 = INFINITY;
        int i0 = 0x7fffffff, i1 = 0x7fffffff;
        for (int it = 0; it < ntiles; it++) {
            const int acc = it & 1;
            mbar_wait(&acc_full[acc], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int col_base = (tile_lo + it) * TC_N;
            // two register buffers: the tcgen05.ld of the next 32 columns is in flight while the
            // current 32 are scanned
            const uint32_t t0 = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * TC_N);
            constexpr int CH = TC_N / 32 / TC_EPI_HALVES;            // 32-column chunks per warp and tile
            const int c_lo = half * CH;
            uint32_t va[32], vb[32];
            tmem_ld32_issue(t0 + (uint32_t)(c_lo * 32), va);
            tmem_ld_wait();
#pragma unroll 1
            for (int c = c_lo; c < c_lo + CH; c += 2) {
                tmem_ld32_issue(t0 + (uint32_t)((c + 1) * 32), vb);
                top2_chunk(va, t0 + (uint32_t)(c * 32), col_base + c * 32, m0, i0, m1, i1);
                tmem_ld_wait();
                if (c + 2 < c_lo + CH) tmem_ld32_issue(t0 + (uint32_t)((c + 2) * 32), va);
                top2_chunk(vb, t0 + (uint32_t)((c + 1) * 32), col_base + (c + 1) * 32, m0, i0, m1, i1);
                tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
        }
        const int qrow = qtile * TC_M + row;
        if (qrow < nq) {
            Cand* o = partial + ((size_t)(split * TC_EPI_HALVES + half) * nq + qrow) * 2;
            o[0].d = m0; o[0].idx = i0;
            o[1].d = m1; o[1].idx = i1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        const uint32_t ncols = 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// The same contraction as a PERSISTENT kernel: one CTA per SM for the whole launch, balanced to one tile.
// Why not more, smaller CTAs: every additional (query tile, train range) costs 13-25 us (15-28 tile times) however short
// the range is -- measured with `matcher_splits` (profiles/r02y_match_splits_*.jsonl: 50 000 x 50 000 with 1 / 2 / 3 / 4 / 6
// splits 0.62 / 0.71 / 0.71 / 0.79 / 0.88 ms, in this kernel and in match_tc_kernel alike: not CTA set-up).  The streaming
// top-2 starts cold (until a row's running second-best is small, nearly every chunk of 32 columns holds a value that beats it
// for one of the warp's 32 rows: ~ 64 ln(n / 64) slow-path columns per warp over n columns), and the first train tiles of a
// range miss L2; which of the two it is was not settled -- inserting cold chunks from the registers instead of re-reading the
// hit columns from TMEM was slower at every size (profiles/r02z_match_cold_list_insertion_dropped.jsonl).  Either way
// splitting the train set for balance never paid, and one CTA per query tile leaves the last wave partly empty (157 query
// tiles on 148 SMs: two waves for 1.06 waves of work).  Here the (query tile, train tile) pairs are
// laid out query-tile-major and cut into gridDim.x EQUAL contiguous ranges: a CTA contracts the tail of one query tile's train
// sequence, whole query tiles, and the head of another ("pieces"), with ~ (query tiles / SMs) + 1 cold starts.  TMEM,
// barriers and pipelines are set up once; the train-tile ring and the two accumulators run across piece boundaries (global
// tile counter), the query tile is double-buffered and the next piece's is fetched while the current one is contracted.
// The pieces of a query tile belong to consecutive CTAs and fill list slots 0, 1, ...; the CTA of the last piece marks the
// slots above as empty.  Same arithmetic, same result bits.  (Measured and dropped: whole query tiles first, all CTAs walking
// the train set in step, and ranges only for the remainder -- every SM asks L2 for the same train tile at the same moment:
// 8.0 ms against 7.4 at 200 000 x 200 000, 1.95 against 1.91 at 100 000; inserting the hit columns from the registers,
// one vote per group of four columns, instead of re-reading them from TMEM: 5-15 % slower at every size.)
constexpr int TCP_SMEM = 2 * TC_A_BYTES + TC_STAGES * TC_B_BYTES + 256;

__global__ void __launch_bounds__(TC_THREADS, 1)
match_tc_persistent_kernel(const uint8_t* __restrict__ qa, const uint8_t* __restrict__ tb, int nq, int q_tiles, int nt_tiles,
                           int lists_per_half, Cand* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                  // [2]
    uint8_t* sB = smem + 2 * TC_A_BYTES;                 // [TC_STAGES]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * TC_A_BYTES + TC_STAGES * TC_B_BYTES);
    uint64_t* a_full = bars + 0;                         // [2]
    uint64_t* a_empty = bars + 2;                        // [2]
    uint64_t* b_full = bars + 4;                         // [TC_STAGES]
    uint64_t* b_empty = bars + 4 + TC_STAGES;
    uint64_t* acc_full = bars + 4 + 2 * TC_STAGES;       // [2]
    uint64_t* acc_empty = bars + 6 + 2 * TC_STAGES;      // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 8 + 2 * TC_STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int a = 0; a < 2; a++) { mbar_init(&a_full[a], 1); mbar_init(&a_empty[a], 1); }
        for (int st = 0; st < TC_STAGES; st++) { mbar_init(&b_full[st], 1); mbar_init(&b_empty[st], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 32 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        const uint32_t ncols = 512;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        // ===== producer: query tiles into the two A buffers (one piece ahead), train tiles through the ring
        if (lane == 0) {
            int ka = 0, gi = 0;                          // query tiles issued, train tiles issued
            auto issue_a = [&](int qtile) {
                const int ab = ka & 1;
                mbar_wait(&a_empty[ab], (((uint32_t)(ka >> 1)) & 1u) ^ 1u);
                mbar_expect_tx(&a_full[ab], TC_A_BYTES);
                bulk_g2s(sA + (size_t)ab * TC_A_BYTES, qa + (size_t)qtile * TC_A_BYTES, TC_A_BYTES, &a_full[ab]);
                ka++;
            };
            TcpPieces it_p(q_tiles, nt_tiles, (int)gridDim.x, (int)blockIdx.x);
            int qtile, tile_lo, ntiles, slot;
            bool have = it_p.next(qtile, tile_lo, ntiles, slot);
            if (have) issue_a(qtile);
            while (have) {
                int q_n, lo_n, n_n, slot_n;
                const bool more = it_p.next(q_n, lo_n, n_n, slot_n);
                // the next query tile goes out once train tile 2 of this piece may be issued: by then the MMAs of the previous
                // piece have completed (its stages came back), so the A buffer it used is free and the wait returns at once
                const int a_at = ntiles > 2 ? 2 : ntiles - 1;
                for (int it = 0; it < ntiles; it++, gi++) {
                    const int st = gi % TC_STAGES;
                    const uint32_t ph = (uint32_t)(gi / TC_STAGES) & 1u;
                    mbar_wait(&b_empty[st], ph ^ 1u);
                    mbar_expect_tx(&b_full[st], TC_B_BYTES);
                    bulk_g2s(sB + (size_t)st * TC_B_BYTES, tb + (size_t)(tile_lo + it) * TC_B_BYTES, TC_B_BYTES, &b_full[st]);
                    if (it == a_at && more) issue_a(q_n);
                }
                have = more; qtile = q_n; tile_lo = lo_n; ntiles = n_n; slot = slot_n;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
            int ka = 0, gi = 0;
            TcpPieces it_p(q_tiles, nt_tiles, (int)gridDim.x, (int)blockIdx.x);
            int qtile, tile_lo, ntiles, slot;
            for (; it_p.next(qtile, tile_lo, ntiles, slot); ka++) {
                const int ab = ka & 1;
                mbar_wait(&a_full[ab], ((uint32_t)(ka >> 1)) & 1u);
                const uint32_t a_addr = s32(sA + (size_t)ab * TC_A_BYTES);
                for (int it = 0; it < ntiles; it++, gi++) {
                    const int st = gi % TC_STAGES, acc = gi & 1;
                    mbar_wait(&b_full[st], (uint32_t)(gi / TC_STAGES) & 1u);
                    mbar_wait(&acc_empty[acc], ((uint32_t)(gi >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t b_addr = s32(sB + (size_t)st * TC_B_BYTES);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TC_N);
#pragma unroll
                    for (int k = 0; k < TC_KCHUNKS / 2; k++)
                        umma_bf16(d_tmem, umma_smem_desc(a_addr + k * 256), umma_smem_desc(b_addr + k * 256), idesc, k > 0);
                    umma_commit(&b_empty[st]);
                    umma_commit(&acc_full[acc]);
                }
                umma_commit(&a_empty[ab]);               // the query tile buffer is free once these MMAs have read it
            }
        }
    } else {
        // ===== epilogue (as in match_tc_kernel), one partial list per piece
        const int lane_grp = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = lane_grp * 32 + lane;
        int gi = 0;
        TcpPieces it_p(q_tiles, nt_tiles, (int)gridDim.x, (int)blockIdx.x);
        int qtile, tile_lo, ntiles, slot;
        while (it_p.next(qtile, tile_lo, ntiles, slot)) {
            float m0 = INFINITY, m1 = INFINITY;
            int i0 = 0x7fffffff, i1 = 0x7fffffff;
            for (int it = 0; it < ntiles; it++, gi++) {
                const int acc = gi & 1;
                mbar_wait(&acc_full[acc], (uint32_t)(gi >> 1) & 1u);
                tc_fence_after();
                const int col_base = (tile_lo + it) * TC_N;
                const uint32_t t0 = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * TC_N);
                constexpr int CH = TC_N / 32 / TC_EPI_HALVES;
                const int c_lo = half * CH;
                uint32_t va[32], vb[32];
                tmem_ld32_issue(t0 + (uint32_t)(c_lo * 32), va);
                tmem_ld_wait();
#pragma unroll 1
                for (int c = c_lo; c < c_lo + CH; c += 2) {
                    tmem_ld32_issue(t0 + (uint32_t)((c + 1) * 32), vb);
                    top2_chunk(va, t0 + (uint32_t)(c * 32), col_base + c * 32, m0, i0, m1, i1);
                    tmem_ld_wait();
                    if (c + 2 < c_lo + CH) tmem_ld32_issue(t0 + (uint32_t)((c + 2) * 32), va);
                    top2_chunk(vb, t0 + (uint32_t)((c + 1) * 32), col_base + (c + 1) * 32, m0, i0, m1, i1);
                    tmem_ld_wait();
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[acc]);
            }
            const int qrow = qtile * TC_M + row;
            if (qrow < nq) {
                Cand* o = partial + ((size_t)(slot * TC_EPI_HALVES + half) * nq + qrow) * 2;
                o[0].d = m0; o[0].idx = i0;
                o[1].d = m1; o[1].idx = i1;
                if (tile_lo + ntiles == nt_tiles) {      // last piece of the query tile: the list slots above stay empty
                    for (int sl = slot + 1; sl < lists_per_half; sl++) {
                        Cand* e = partial + ((size_t)(sl * TC_EPI_HALVES + half) * nq + qrow) * 2;
                        e[0].d = INFINITY; e[0].idx = 0x7fffffff;
                        e[1].d = INFINITY; e[1].idx = 0x7fffffff;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        const uint32_t ncols = 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}


// ------------------------------------------------------------------------------------------
// Tensor-core path for REAL-VALUED float descriptors (SURF, RootSIFT, learned descriptors; dim <= 128):
// filter on the tensor cores, decide exactly on CUDA cores.
//   filter   every value is split x = hi + lo with hi, lo in bf16 (|x - hi - lo| <= 2^-18 |x|); the
//            accumulator collects (-2q)_hi.t_hi + (-2q)_hi.t_lo + (-2q)_lo.t_hi + |t|^2 (three bf16 pieces),
//            i.e. d^2 - |q|^2 up to E = 2^-11 (|q||t| + |t|^2) (a deliberately loose bound: the split leaves
//            6 x 2^-18 |q||t|, the rest covers the fp32 accumulation inside the tensor core).  Each query
//            row keeps its FOUR smallest values per train split.
//   decide   the candidates are re-evaluated with the arithmetic of match_f32_kernel (ascending
//            dimensions, separately rounded multiply and add), the two smallest by (d, index) win.  A train
//            descriptor that is not a candidate has approximate value >= the smallest 4th value L of any
//            split, hence exact d^2 >= L + |q|^2 - E: if the exact second-best is below that, the result
//            is the exact brute-force answer; otherwise the query is handed to match_f32_kernel.
// M128 N256 K16 x (1 + 3 ceil(dim/16)) per tile; operands 34 K-chunks per row (16 hi, 16 lo, 2 extras).
// ------------------------------------------------------------------------------------------
constexpr int SP_M = 128;
constexpr int SP_TOPK = 4;
// Row layout: [hi: dpad/8 chunks | lo: dpad/8 chunks | extras: 2 chunks] of 16 bytes, dpad = dim rounded up to 16.
// dim 128: 34 chunks, 136 KB train tile -> one stage next to the 68 KB query tile; dim <= 80: two stages fit.
struct SpLayout {
    int n_tile;         // train rows per tile: 256, or 128 when that is what allows two stages
    int half_chunks;    // dpad / 8
    int kchunks;        // 2 * half_chunks + 2
    int group_bytes;    // kchunks * 128 (8 rows)
    int a_bytes, b_bytes, stages, smem;
    int a_bufs, smem_persistent;
};
inline SpLayout sp_layout(int dim, size_t smem_optin, int n_tile_option) {
    SpLayout L;
    const int dpad = (dim + 15) / 16 * 16;
    L.half_chunks = dpad / 8;
    L.kchunks = 2 * L.half_chunks + 2;
    L.group_bytes = L.kchunks * 128;
    L.a_bytes = (SP_M / 8) * L.group_bytes;
    L.n_tile = 256;
    L.b_bytes = (L.n_tile / 8) * L.group_bytes;
    L.stages = (size_t)L.a_bytes + 2 * (size_t)L.b_bytes + 256 <= smem_optin ? 2 : 1;
    if (L.stages == 1 && n_tile_option == 128) {            // two stages of 128 rows instead of one of 256
        L.n_tile = 128;
        L.b_bytes = (L.n_tile / 8) * L.group_bytes;
        L.stages = 2;
    }
    L.smem = L.a_bytes + L.stages * L.b_bytes + 256;
    // persistent kernel: a second query-tile buffer where it fits (dim <= 80), else the query tile of a piece is fetched
    // when the previous piece has been contracted
    L.a_bufs = (size_t)2 * L.a_bytes + (size_t)L.stages * L.b_bytes + 256 <= smem_optin ? 2 : 1;
    L.smem_persistent = L.a_bufs * L.a_bytes + L.stages * L.b_bytes + 256;
    return L;
}

struct Cand4 { float d[SP_TOPK]; int idx[SP_TOPK]; };

__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// One warp per row: re-tile into [hi chunks 0..15 | lo chunks 16..31 | extras 32..33]; norms[row] = |x|^2.
__global__ void __launch_bounds__(256)
sp_prep_kernel(const float* __restrict__ src, int n, int n_pad, int dim, int is_query, int half_chunks, uint8_t* __restrict__ dst,
               float* __restrict__ norms, unsigned* __restrict__ max_norm_bits) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n_pad) return;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < n && lane * 4 < dim) {
        const float4 x = *reinterpret_cast<const float4*>(src + (size_t)row * dim + lane * 4);
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    }
    float ss = v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const int group_bytes = (2 * half_chunks + 2) * 128;
    uint8_t* g = dst + (size_t)(row >> 3) * group_bytes + (row & 7) * 16;
    const float sc = is_query ? -2.f : 1.f;
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; k++) split_bf16(sc * v[k], h[k], l[k]);
    uint2 ph, pl;
    ph.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
    if ((lane >> 1) < half_chunks) {
        *reinterpret_cast<uint2*>(g + (lane >> 1) * 128 + (lane & 1) * 8) = ph;
        *reinterpret_cast<uint2*>(g + (half_chunks + (lane >> 1)) * 128 + (lane & 1) * 8) = pl;
    }
    if (lane == 0) {
        float ex[16];
#pragma unroll
        for (int k = 0; k < 16; k++) ex[k] = 0.f;
        if (is_query) {
            ex[0] = ex[1] = ex[2] = 1.f;
        } else {
            // |t|^2 in three bf16 pieces; padded rows get a huge norm so that they never enter a top-4
            const float nn = row < n ? ss : 1.0e30f;
            __nv_bfloat16 n0 = __float2bfloat16_rn(nn);
            const float r1 = nn - __bfloat162float(n0);
            __nv_bfloat16 n1 = __float2bfloat16_rn(r1);
            const float r2 = r1 - __bfloat162float(n1);
            ex[0] = __bfloat162float(n0); ex[1] = __bfloat162float(n1); ex[2] = __bfloat162float(__float2bfloat16_rn(r2));
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                __nv_bfloat162 p = __floats2bfloat162_rn(ex[c * 8 + 2 * k], ex[c * 8 + 2 * k + 1]);
                w[k] = *reinterpret_cast<uint32_t*>(&p);
            }
            *reinterpret_cast<uint4*>(g + (2 * half_chunks + c) * 128) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        if (row < n) {
            norms[row] = ss;
            if (ss == ss && ss < 3.0e38f) atomicMax(max_norm_bits, __float_as_uint(ss));
            else atomicMax(max_norm_bits, 0x7f800000u);      // NaN / inf in the data: the exact path decides
        }
    }
}

__device__ __forceinline__ uint64_t umma_smem_desc_g(uint32_t saddr, uint32_t group_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((128u >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((group_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return v;
}

// Running top-4 (ascending) of one query row over 32 accumulator columns.  The common case is a min tree and one
// vote; columns that beat some row's 4th value are few, and they are handled by a COMPACT loop that re-reads the
// column from TMEM (a 32-way unrolled insertion sequence per chunk is several thousand instructions: the kernel
// then stalls on instruction fetch, `no_instruction` 4.5 cycles per issue in profiles/r01f_match_sp_*).
__device__ __forceinline__ void top4_chunk(const uint32_t (&v)[32], uint32_t taddr, int col0, float (&m)[SP_TOPK], int (&ix)[SP_TOPK]) {
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; j++)
        p[j] = fminf(fminf(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])),
                     fminf(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])));
    const float gmin = fminf(fminf(fminf(p[0], p[1]), fminf(p[2], p[3])), fminf(fminf(p[4], p[5]), fminf(p[6], p[7])));
    if (__any_sync(0xffffffffu, gmin < m[3])) {
        unsigned mask = 0u;
#pragma unroll
        for (int k = 0; k < 32; k++) mask |= (__uint_as_float(v[k]) < m[3]) ? (1u << k) : 0u;
        unsigned um = __reduce_or_sync(0xffffffffu, mask);
        while (um) {                                        // warp-uniform: ascending columns, so ties keep the lower index
            const int k = __ffs(um) - 1;
            um &= um - 1u;
            const float d = __uint_as_float(tmem_ld1(taddr + (uint32_t)k));
            if (d < m[3]) {
                const int j = col0 + k;
                if (d < m[0]) { m[3] = m[2]; ix[3] = ix[2]; m[2] = m[1]; ix[2] = ix[1]; m[1] = m[0]; ix[1] = ix[0]; m[0] = d; ix[0] = j; }
                else if (d < m[1]) { m[3] = m[2]; ix[3] = ix[2]; m[2] = m[1]; ix[2] = ix[1]; m[1] = d; ix[1] = j; }
                else if (d < m[2]) { m[3] = m[2]; ix[3] = ix[2]; m[2] = d; ix[2] = j; }
                else { m[3] = d; ix[3] = j; }
            }
        }
    }
}

constexpr int SP_EPI_HALVES = 1;                    // measured: a second epilogue warp per lane quarter does not pay at N = 128
constexpr int SP_THREADS = 64 + 128 * SP_EPI_HALVES;
template <int SP_N>
__global__ void __launch_bounds__(SP_THREADS, 1)
match_sp_kernel(const uint8_t* __restrict__ qa, const uint8_t* __restrict__ tb, int nq, int nt_tiles,
                int tiles_per_split, int ksteps, const SpLayout lay, Cand4* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int SP_A_BYTES = lay.a_bytes, SP_B_BYTES = lay.b_bytes, SP_STAGES = lay.stages, SP_GROUP_BYTES = lay.group_bytes;
    uint8_t* sA = smem;
    uint8_t* sB = smem + SP_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SP_A_BYTES + SP_STAGES * SP_B_BYTES);
    uint64_t* a_full = bars + 0;
    uint64_t* b_full = bars + 1;            // [2]
    uint64_t* b_empty = bars + 3;           // [2]
    uint64_t* acc_full = bars + 5;          // [2]
    uint64_t* acc_empty = bars + 7;         // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qtile = blockIdx.x, split = blockIdx.y;
    const int tile_lo = split * tiles_per_split;
    const int tile_hi = min(nt_tiles, tile_lo + tiles_per_split);
    const int ntiles = tile_hi - tile_lo;

    if (threadIdx.x == 0) {
        mbar_init(a_full, 1);
        for (int s = 0; s < SP_STAGES; s++) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128 * SP_EPI_HALVES); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        const uint32_t ncols = 2 * SP_N;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(a_full, SP_A_BYTES);
            bulk_g2s(sA, qa + (size_t)qtile * SP_A_BYTES, SP_A_BYTES, a_full);
            for (int it = 0; it < ntiles; it++) {
                const int s = it % SP_STAGES;
                const uint32_t ph = (uint32_t)(it / SP_STAGES) & 1u;
                mbar_wait(&b_empty[s], ph ^ 1u);
                mbar_expect_tx(&b_full[s], SP_B_BYTES);
                bulk_g2s(sB + (size_t)s * SP_B_BYTES, tb + (size_t)(tile_lo + it) * SP_B_BYTES, SP_B_BYTES, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SP_N >> 3) << 17) | ((uint32_t)(SP_M >> 4) << 24);
            mbar_wait(a_full, 0);
            const uint32_t a_addr = s32(sA);
            for (int it = 0; it < ntiles; it++) {
                const int s = it % SP_STAGES, acc = it & 1;
                mbar_wait(&b_full[s], (uint32_t)(it / SP_STAGES) & 1u);
                mbar_wait(&acc_empty[acc], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t b_addr = s32(sB + (size_t)s * SP_B_BYTES);
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SP_N);
                // extras first (|t|^2), then per K16 step hi.hi, hi.lo, lo.hi
                const uint32_t ex_off = (uint32_t)(2 * lay.half_chunks) * 128u, lo_off = (uint32_t)lay.half_chunks * 128u;
                umma_bf16(d_tmem, umma_smem_desc_g(a_addr + ex_off, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + ex_off, SP_GROUP_BYTES), idesc, 0);
#pragma unroll 2
                for (int k = 0; k < ksteps; k++) {          // K16 steps: ceil(dim / 16)
                    const uint32_t hi = k * 256, lo = lo_off + k * 256;
                    umma_bf16(d_tmem, umma_smem_desc_g(a_addr + hi, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + hi, SP_GROUP_BYTES), idesc, 1);
                    umma_bf16(d_tmem, umma_smem_desc_g(a_addr + hi, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + lo, SP_GROUP_BYTES), idesc, 1);
                    umma_bf16(d_tmem, umma_smem_desc_g(a_addr + lo, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + hi, SP_GROUP_BYTES), idesc, 1);
                }
                umma_commit(&b_empty[s]);
                umma_commit(&acc_full[acc]);
            }
        }
    } else {
        const int lane_grp = warp & 3;
        const int half = (warp - 2) >> 2;               // two epilogue warps per lane quarter, half the columns each
        const int row = lane_grp * 32 + lane;
        float m[SP_TOPK];
        int ix[SP_TOPK];
#pragma unroll
        for (int k = 0; k < SP_TOPK; k++) { m[k] = INFINITY; ix[k] = 0x7fffffff; }
        for (int it = 0; it < ntiles; it++) {
            const int acc = it & 1;
            mbar_wait(&acc_full[acc], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int col_base = (tile_lo + it) * SP_N;
            const uint32_t t0 = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * SP_N);
            constexpr int CH = SP_N / 32 / SP_EPI_HALVES;
            const int c_lo = half * CH;
            uint32_t va[32], vb[32];
            tmem_ld32_issue(t0 + (uint32_t)(c_lo * 32), va);
            tmem_ld_wait();
#pragma unroll 1
            for (int c = c_lo; c < c_lo + CH; c += 2) {
                tmem_ld32_issue(t0 + (uint32_t)((c + 1) * 32), vb);
                top4_chunk(va, t0 + (uint32_t)(c * 32), col_base + c * 32, m, ix);
                tmem_ld_wait();
                if (c + 2 < c_lo + CH) tmem_ld32_issue(t0 + (uint32_t)((c + 2) * 32), va);
                top4_chunk(vb, t0 + (uint32_t)((c + 1) * 32), col_base + (c + 1) * 32, m, ix);
                tmem_ld_wait();
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[acc]);
        }
        const int qrow = qtile * SP_M + row;
        if (qrow < nq) {
            Cand4* o = partial + (size_t)(split * SP_EPI_HALVES + half) * nq + qrow;
#pragma unroll
            for (int k = 0; k < SP_TOPK; k++) { o->d[k] = m[k]; o->idx[k] = ix[k]; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        const uint32_t ncols = 2 * SP_N;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// match_sp_kernel as a persistent kernel over equal ranges of the (query tile, train tile) sequence: see
// match_tc_persistent_kernel (the top-4 of the filter starts even colder than a top-2).  Two query-tile buffers where shared
// memory allows (dim <= 80); with one (dim 128: 68 KB query tile + two 68 KB train stages) the query tile of a piece goes
// out right after the piece's first train tile, as soon as the MMAs of the previous piece have read the buffer.
template <int SP_N>
__global__ void __launch_bounds__(SP_THREADS, 1)
match_sp_persistent_kernel(const uint8_t* __restrict__ qa, const uint8_t* __restrict__ tb, int nq, int q_tiles, int nt_tiles,
                           int lists_per_half, int ksteps, const SpLayout lay, Cand4* __restrict__ partial) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int SP_A_BYTES = lay.a_bytes, SP_B_BYTES = lay.b_bytes, SP_STAGES = lay.stages, SP_GROUP_BYTES = lay.group_bytes;
    const int NBUF = lay.a_bufs;
    uint8_t* sA = smem;
    uint8_t* sB = smem + (size_t)NBUF * SP_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NBUF * SP_A_BYTES + (size_t)SP_STAGES * SP_B_BYTES);
    uint64_t* a_full = bars + 0;            // [2]
    uint64_t* a_empty = bars + 2;           // [2]
    uint64_t* b_full = bars + 4;            // [2]
    uint64_t* b_empty = bars + 6;           // [2]
    uint64_t* acc_full = bars + 8;          // [2]
    uint64_t* acc_empty = bars + 10;        // [2]
    uint32_t* tmem_base_s = reinterpret_cast<uint32_t*>(bars + 12);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int a = 0; a < 2; a++) { mbar_init(&a_full[a], 1); mbar_init(&a_empty[a], 1); }
        for (int st = 0; st < 2; st++) { mbar_init(&b_full[st], 1); mbar_init(&b_empty[st], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128 * SP_EPI_HALVES); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        const uint32_t ncols = 2 * SP_N;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_base_s)), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            int ka = 0, gi = 0;
            auto issue_a = [&](int qtile) {
                const int ab = ka % NBUF;
                mbar_wait(&a_empty[ab], (((uint32_t)(ka / NBUF)) & 1u) ^ 1u);
                mbar_expect_tx(&a_full[ab], SP_A_BYTES);
                bulk_g2s(sA + (size_t)ab * SP_A_BYTES, qa + (size_t)qtile * SP_A_BYTES, SP_A_BYTES, &a_full[ab]);
                ka++;
            };
            TcpPieces it_p(q_tiles, nt_tiles, (int)gridDim.x, (int)blockIdx.x);
            int qtile, tile_lo, ntiles, slot;
            bool have = it_p.next(qtile, tile_lo, ntiles, slot);
            if (have && NBUF == 2) issue_a(qtile);
            while (have) {
                int q_n, lo_n, n_n, slot_n;
                const bool more = it_p.next(q_n, lo_n, n_n, slot_n);
                const int a_at = ntiles > 2 ? 2 : ntiles - 1;
                for (int it = 0; it < ntiles; it++, gi++) {
                    const int st = gi % SP_STAGES;
                    const uint32_t ph = (uint32_t)(gi / SP_STAGES) & 1u;
                    mbar_wait(&b_empty[st], ph ^ 1u);
                    mbar_expect_tx(&b_full[st], SP_B_BYTES);
                    bulk_g2s(sB + (size_t)st * SP_B_BYTES, tb + (size_t)(tile_lo + it) * SP_B_BYTES, SP_B_BYTES, &b_full[st]);
                    if (NBUF == 2) { if (it == a_at && more) issue_a(q_n); }
                    else if (it == 0) issue_a(qtile);
                }
                have = more; qtile = q_n; tile_lo = lo_n; ntiles = n_n; slot = slot_n;
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SP_N >> 3) << 17) | ((uint32_t)(SP_M >> 4) << 24);
            const uint32_t ex_off = (uint32_t)(2 * lay.half_chunks) * 128u, lo_off = (uint32_t)lay.half_chunks * 128u;
            int ka = 0, gi = 0;
            TcpPieces it_p(q_tiles, nt_tiles, (int)gridDim.x, (int)blockIdx.x);
            int qtile, tile_lo, ntiles, slot;
            for (; it_p.next(qtile, tile_lo, ntiles, slot); ka++) {
                const int ab = ka % NBUF;
                mbar_wait(&a_full[ab], ((uint32_t)(ka / NBUF)) & 1u);
                const uint32_t a_addr = s32(sA + (size_t)ab * SP_A_BYTES);
                for (int it = 0; it < ntiles; it++, gi++) {
                    const int st = gi % SP_STAGES, acc = gi & 1;
                    mbar_wait(&b_full[st], (uint32_t)(gi / SP_STAGES) & 1u);
                    mbar_wait(&acc_empty[acc], ((uint32_t)(gi >> 1) & 1u) ^ 1u);
                    tc_fence_after();
                    const uint32_t b_addr = s32(sB + (size_t)st * SP_B_BYTES);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(acc * SP_N);
                    umma_bf16(d_tmem, umma_smem_desc_g(a_addr + ex_off, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + ex_off, SP_GROUP_BYTES), idesc, 0);
#pragma unroll 2
                    for (int k = 0; k < ksteps; k++) {
                        const uint32_t hi = k * 256, lo = lo_off + k * 256;
                        umma_bf16(d_tmem, umma_smem_desc_g(a_addr + hi, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + hi, SP_GROUP_BYTES), idesc, 1);
                        umma_bf16(d_tmem, umma_smem_desc_g(a_addr + hi, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + lo, SP_GROUP_BYTES), idesc, 1);
                        umma_bf16(d_tmem, umma_smem_desc_g(a_addr + lo, SP_GROUP_BYTES), umma_smem_desc_g(b_addr + hi, SP_GROUP_BYTES), idesc, 1);
                    }
                    umma_commit(&b_empty[st]);
                    umma_commit(&acc_full[acc]);
                }
                umma_commit(&a_empty[ab]);
            }
        }
    } else {
        const int lane_grp = warp & 3;
        const int half = (warp - 2) >> 2;
        const int row = lane_grp * 32 + lane;
        int gi = 0;
        TcpPieces it_p(q_tiles, nt_tiles, (int)gridDim.x, (int)blockIdx.x);
        int qtile, tile_lo, ntiles, slot;
        while (it_p.next(qtile, tile_lo, ntiles, slot)) {
            float m[SP_TOPK];
            int ix[SP_TOPK];
#pragma unroll
            for (int k = 0; k < SP_TOPK; k++) { m[k] = INFINITY; ix[k] = 0x7fffffff; }
            for (int it = 0; it < ntiles; it++, gi++) {
                const int acc = gi & 1;
                mbar_wait(&acc_full[acc], (uint32_t)(gi >> 1) & 1u);
                tc_fence_after();
                const int col_base = (tile_lo + it) * SP_N;
                const uint32_t t0 = tmem_base + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(acc * SP_N);
                constexpr int CH = SP_N / 32 / SP_EPI_HALVES;
                const int c_lo = half * CH;
                uint32_t va[32], vb[32];
                tmem_ld32_issue(t0 + (uint32_t)(c_lo * 32), va);
                tmem_ld_wait();
#pragma unroll 1
                for (int c = c_lo; c < c_lo + CH; c += 2) {
                    tmem_ld32_issue(t0 + (uint32_t)((c + 1) * 32), vb);
                    top4_chunk(va, t0 + (uint32_t)(c * 32), col_base + c * 32, m, ix);
                    tmem_ld_wait();
                    if (c + 2 < c_lo + CH) tmem_ld32_issue(t0 + (uint32_t)((c + 2) * 32), va);
                    top4_chunk(vb, t0 + (uint32_t)((c + 1) * 32), col_base + (c + 1) * 32, m, ix);
                    tmem_ld_wait();
                }
                tc_fence_before();
                mbar_arrive(&acc_empty[acc]);
            }
            const int qrow = qtile * SP_M + row;
            if (qrow < nq) {
                Cand4* o = partial + (size_t)(slot * SP_EPI_HALVES + half) * nq + qrow;
#pragma unroll
                for (int k = 0; k < SP_TOPK; k++) { o->d[k] = m[k]; o->idx[k] = ix[k]; }
                if (tile_lo + ntiles == nt_tiles) {      // last piece of the query tile: the list slots above stay empty
                    for (int sl = slot + 1; sl < lists_per_half; sl++) {
                        Cand4* e = partial + (size_t)(sl * SP_EPI_HALVES + half) * nq + qrow;
#pragma unroll
                        for (int k = 0; k < SP_TOPK; k++) { e->d[k] = INFINITY; e->idx[k] = 0x7fffffff; }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        const uint32_t ncols = 2 * SP_N;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}


// Exact decision for one query per thread: re-evaluate its candidates with match_f32_kernel's arithmetic, keep
// the two smallest by (d, index), and prove that no other train descriptor can beat the second (else: flag).
__global__ void sp_refine_kernel(const float* __restrict__ q, int nq, const float* __restrict__ t, int nt, int dim,
                                 const Cand4* __restrict__ partial, int nsplit, const float* __restrict__ qnorm,
                                 const unsigned* __restrict__ max_tnorm_bits, int32_t* __restrict__ idx, float* __restrict__ dist,
                                 int* __restrict__ n_flagged, int32_t* __restrict__ flagged) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const float* qi = q + (size_t)i * dim;
    float m0 = INFINITY, m1 = INFINITY, low = INFINITY;
    int j0 = 0x7fffffff, j1 = 0x7fffffff;
    for (int s = 0; s < nsplit; s++) {
        const Cand4 c = partial[(size_t)s * nq + i];
        low = fminf(low, c.d[SP_TOPK - 1]);           // +inf when the split holds fewer than four descriptors
#pragma unroll
        for (int k = 0; k < SP_TOPK; k++) {
            const int j = c.idx[k];
            if (j < 0 || j >= nt) continue;
            const float* tj = t + (size_t)j * dim;
            float acc = 0.f;
            for (int d = 0; d < dim; d += 4) {
                const float4 a = *reinterpret_cast<const float4*>(qi + d);
                const float4 b = *reinterpret_cast<const float4*>(tj + d);
                float e = __fsub_rn(a.x, b.x); acc = __fadd_rn(acc, __fmul_rn(e, e));
                e = __fsub_rn(a.y, b.y); acc = __fadd_rn(acc, __fmul_rn(e, e));
                e = __fsub_rn(a.z, b.z); acc = __fadd_rn(acc, __fmul_rn(e, e));
                e = __fsub_rn(a.w, b.w); acc = __fadd_rn(acc, __fmul_rn(e, e));
            }
            top2_insert(acc, j, m0, j0, m1, j1);
        }
    }
    // The filter value of every non-candidate is >= low, so its exact d^2 is >= low + |q|^2 - E with
    //   E = 2^-11 (|q||t|max + |t|max^2): the bf16 split leaves 6 x 2^-18 |q||t| and 2^-24 |t|^2, a few hundred fp32
    //       additions inside the tensor core at most ~1e-4 (2|q||t| + |t|^2) even if they truncated;
    // and what match_f32_kernel would compute for it (the value the result is defined by) is within
    //   2 dim 2^-24 d^2 <= 1.6e-5 (|q| + |t|max)^2 of the exact d^2 (sequential fp32 sum of dim <= 128 terms).
    const float qn = qnorm[i], tmax = __uint_as_float(*max_tnorm_bits);
    const float qt = sqrtf(qn) * sqrtf(tmax);
    const float E = 4.8828125e-4f * (qt + tmax) + 1.6e-5f * (qn + 2.0f * qt + tmax);
    const float bound = (low + qn) - E - 1.0e-6f * (fabsf(low) + qn);    // last term: rounding of this expression itself
    const bool safe = j1 < nt && m1 < bound;      // low = +inf: every train descriptor was a candidate
    if (safe) {
        idx[2 * i] = j0; idx[2 * i + 1] = j1;
        dist[2 * i] = sqrtf(m0); dist[2 * i + 1] = sqrtf(m1);
    } else {
        idx[2 * i] = -2;                              // decided by the exact path
        flagged[atomicAdd(n_flagged, 1)] = i;
    }
}

__global__ void sp_gather_rows_kernel(const float* __restrict__ q, int dim, const int32_t* __restrict__ rows, int n,
                                      float* __restrict__ out) {
    const int r = blockIdx.x;
    if (r >= n) return;
    for (int d = threadIdx.x; d < dim; d += blockDim.x) out[(size_t)r * dim + d] = q[(size_t)rows[r] * dim + d];
}
__global__ void sp_scatter_kernel(const int32_t* __restrict__ rows, int n, const int32_t* __restrict__ sidx, const float* __restrict__ sdist,
                                  int32_t* __restrict__ idx, float* __restrict__ dist) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int i = rows[r];
    idx[2 * i] = sidx[2 * r]; idx[2 * i + 1] = sidx[2 * r + 1];
    dist[2 * i] = sdist[2 * r]; dist[2 * i + 1] = sdist[2 * r + 1];
}

int pick_splits(int work_tiles, int nt_tiles, int sms, int* tiles_per_split) {
    int splits = 1;
    if (work_tiles < 2 * sms) splits = (2 * sms + work_tiles - 1) / work_tiles;
    if (splits > nt_tiles) splits = nt_tiles;
    if (splits < 1) splits = 1;
    *tiles_per_split = (nt_tiles + splits - 1) / splits;
    if (*tiles_per_split < 1) *tiles_per_split = 1;
    const int eff = (nt_tiles + *tiles_per_split - 1) / *tiles_per_split;
    return eff > 0 ? eff : 1;
}

// The CTAs of the persistent kernels are at different places of the train set at any time, so the whole re-tiled train
// operand has to stay in L2 (126 MB on B200, of which about half serves one die's SMs): real-valued 128-d descriptors at
// 200 000 x 200 000 (109 MB) took 34 ms against 28 ms with one CTA per query tile (whose CTAs walk the train set more or less
// together) -- 170 GB of train tiles through HBM; at 54 MB (100 000) and 58 MB (integer path, 200 000) the persistent kernels
// are ahead.  `matcher_persistent` = 2 forces them.
bool tcp_train_fits_l2(const fm3d_ctx* ctx, size_t train_operand_bytes) {
    return ctx->opt_matcher_persistent == 2 || train_operand_bytes <= ((size_t)64 << 20);
}

// Train splits of the one-CTA-per-item tensor-core kernels: round 1's rule (split only below 2 CTAs per SM), or as many as
// `forced` > 0 says (measurements: every split costs another cold start of the streaming top-2, see match_tc_persistent_kernel).
int pick_splits_waves(int q_tiles, int nt_tiles, int sms, int forced, int* tiles_per_split) {
    if (forced <= 0) return pick_splits(q_tiles, nt_tiles, sms / 2 + 1, tiles_per_split);
    int best = forced > nt_tiles ? nt_tiles : forced;
    if (best < 1) best = 1;
    *tiles_per_split = (nt_tiles + best - 1) / best;
    if (*tiles_per_split < 1) *tiles_per_split = 1;
    const int eff = (nt_tiles + *tiles_per_split - 1) / *tiles_per_split;
    return eff > 0 ? eff : 1;
}

// exact CUDA-core path (K1')
int knn2_f32_generic(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim, int32_t* idx, float* dist) {
    const int sms = ctx->prop.multiProcessorCount;
    const int q_tiles = (nq + FT - 1) / FT, nt_tiles = (nt + FT - 1) / FT;
    int tps = 1;
    const int splits = nt_tiles > 0 ? pick_splits(q_tiles, nt_tiles, sms, &tps) : 1;
    Cand* partial = nullptr;
    if (int rc = fm3d_scratch(ctx, 4, sizeof(Cand) * 2 * (size_t)splits * nq, (void**)&partial)) return rc;
    dim3 grid(q_tiles, splits);
    match_f32_kernel<<<grid, 256, 0, ctx->stream>>>(q, nq, t, nt, dim, tps, partial);
    FM3D_LAUNCH_CHECK(ctx);
    finalize_f32_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(partial, splits, nq, nt, idx, dist);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int knn2_f32_dev(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim, int32_t* idx,
                 float* dist) {
    const int sms = ctx->prop.multiProcessorCount;
    ctx->n_matcher_exact_fallback = 0;
    bool use_tc = ctx->opt_matcher_tensor && dim == TC_DIM && nt >= 1 &&
                  (((uintptr_t)q | (uintptr_t)t) & 15) == 0;
    if (use_tc) {
        const int nq_pad = (nq + TC_M - 1) / TC_M * TC_M, nt_pad = (nt + TC_N - 1) / TC_N * TC_N;
        const size_t ba = (size_t)(nq_pad / 8) * TC_GROUP_BYTES, bb = (size_t)(nt_pad / 8) * TC_GROUP_BYTES;
        uint8_t* ops = nullptr;
        if (int rc = fm3d_scratch(ctx, 3, ba + bb + 256, (void**)&ops)) return rc;
        int* flag = reinterpret_cast<int*>(ops + ba + bb);
        ctx->n_copy++;
        FM3D_CUDA(ctx, cudaMemsetAsync(flag, 0, sizeof(int), ctx->stream));
        tc_prep_kernel<<<nq_pad / 8, 256, 0, ctx->stream>>>(q, nq, nq_pad, 1, ops, flag);
        FM3D_LAUNCH_CHECK(ctx);
        tc_prep_kernel<<<nt_pad / 8, 256, 0, ctx->stream>>>(t, nt, nt_pad, 0, ops + ba, flag);
        FM3D_LAUNCH_CHECK(ctx);
        // The contraction is launched WITHOUT waiting for the verdict of the two prep kernels ("every value is an integer in
        // [0, 255]"): the four kernels run back to back and the flag is read once, after them.  Integer descriptors (OpenCV
        // SIFT, the case this path exists for) never pay a mid-pipeline host round trip; for anything else the speculative
        // result is overwritten by the exact paths below -- and the context remembers the verdict: after a call with
        // real-valued 128-d descriptors (SURF-128, RootSIFT) the next one reads the flag BEFORE the contraction instead of
        // wasting it (7.7 ms at 200 000 x 200 000 next to the 26 ms of the filter), until integer descriptors show up again.
        bool contract = true;
        if (!ctx->matcher_expect_integer) {
            int h_flag = 0;
            if (int rc = fm3d_d2h(ctx, &h_flag, flag, sizeof(int))) return rc;
            FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            contract = h_flag == 0;
            if (contract) ctx->matcher_expect_integer = 1;
        }
        if (contract) {
            const int q_tiles = nq_pad / TC_M, nt_tiles = nt_pad / TC_N;
            int tps = 1;
            // persistent CTAs over equal ranges of the flattened (query tile, train tile) sequence (default), or one CTA per
            // (query tile, train split) with round 1's split rule (`matcher_persistent` = 0; `matcher_splits` forces the splits)
            const bool persistent = ctx->opt_matcher_persistent != 0 && ctx->opt_matcher_splits == 0 && tcp_train_fits_l2(ctx, bb) &&
                                    (size_t)TCP_SMEM <= ctx->prop.sharedMemPerBlockOptin;
            Cand* partial = nullptr;
            int lists = 0;
            if (persistent) {
                int G = 1, pieces = 1;
                tcp_plan(q_tiles, nt_tiles, sms, ctx->opt_matcher_min_tiles, &G, &pieces);
                lists = pieces * TC_EPI_HALVES;
                if (int rc = fm3d_scratch(ctx, 4, sizeof(Cand) * 2 * (size_t)lists * nq, (void**)&partial)) return rc;
                FM3D_CUDA(ctx, cudaFuncSetAttribute(match_tc_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TCP_SMEM));
                match_tc_persistent_kernel<<<G, TC_THREADS, TCP_SMEM, ctx->stream>>>(ops, ops + ba, nq, q_tiles, nt_tiles, pieces, partial);
            } else {
                const int splits = pick_splits_waves(q_tiles, nt_tiles, sms, ctx->opt_matcher_splits, &tps);
                lists = splits * TC_EPI_HALVES;
                if (int rc = fm3d_scratch(ctx, 4, sizeof(Cand) * 2 * (size_t)lists * nq, (void**)&partial)) return rc;
                FM3D_CUDA(ctx, cudaFuncSetAttribute(match_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
                dim3 grid(q_tiles, splits);
                match_tc_kernel<<<grid, TC_THREADS, TC_SMEM, ctx->stream>>>(ops, ops + ba, nq, nt_tiles, tps, partial);
            }
            FM3D_LAUNCH_CHECK(ctx);
            finalize_f32_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(partial, lists, nq, nt, idx, dist);
            FM3D_LAUNCH_CHECK(ctx);
            int h_flag = 0;
            if (int rc = fm3d_d2h(ctx, &h_flag, flag, sizeof(int))) return rc;
            FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (h_flag == 0) return FM3D_OK;
            ctx->matcher_expect_integer = 0;
        }
    }
    // real-valued descriptors: bf16 hi/lo filter on the tensor cores + exact decision (see match_sp_kernel)
    const bool use_sp = ctx->opt_matcher_tensor && dim <= TC_DIM && (dim & 3) == 0 && nt >= 1 &&
                        (((uintptr_t)q | (uintptr_t)t) & 15) == 0 && (long long)nq * (long long)nt >= (1ll << 22);
    if (use_sp) {
        const SpLayout lay = sp_layout(dim, ctx->prop.sharedMemPerBlockOptin, ctx->opt_matcher_sp_tile);
        const int SP_N = lay.n_tile;
        const int nq_pad = (nq + SP_M - 1) / SP_M * SP_M, nt_pad = (nt + SP_N - 1) / SP_N * SP_N;
        const size_t ba = (size_t)(nq_pad / 8) * lay.group_bytes, bb = (size_t)(nt_pad / 8) * lay.group_bytes;
        const size_t bn = (sizeof(float) * ((size_t)nq + nt) + 255) & ~(size_t)255;
        uint8_t* ops = nullptr;
        if (int rc = fm3d_scratch(ctx, 3, ba + bb + bn + 256, (void**)&ops)) return rc;
        float* qnorm = reinterpret_cast<float*>(ops + ba + bb);
        float* tnorm = qnorm + nq;
        unsigned* flags = reinterpret_cast<unsigned*>(ops + ba + bb + bn);   // [0] max |q|^2 bits, [1] max |t|^2 bits, [2] flagged count
        ctx->n_copy++;
        FM3D_CUDA(ctx, cudaMemsetAsync(flags, 0, 16, ctx->stream));
        sp_prep_kernel<<<nq_pad / 8, 256, 0, ctx->stream>>>(q, nq, nq_pad, dim, 1, lay.half_chunks, ops, qnorm, flags);
        FM3D_LAUNCH_CHECK(ctx);
        sp_prep_kernel<<<nt_pad / 8, 256, 0, ctx->stream>>>(t, nt, nt_pad, dim, 0, lay.half_chunks, ops + ba, tnorm, flags + 1);
        FM3D_LAUNCH_CHECK(ctx);
        const int q_tiles = nq_pad / SP_M, nt_tiles = nt_pad / SP_N;
        int tps = 1;
        const bool persistent = ctx->opt_matcher_persistent != 0 && ctx->opt_matcher_splits == 0 && tcp_train_fits_l2(ctx, bb);
        Cand4* partial4 = nullptr;
        int32_t* flagged = nullptr;
        if (int rc = fm3d_scratch(ctx, 5, sizeof(int32_t) * (size_t)nq, (void**)&flagged)) return rc;
        int lists = 0;
        if (persistent) {
            int G = 1, pieces = 1;
            tcp_plan(q_tiles, nt_tiles, sms, ctx->opt_matcher_min_tiles, &G, &pieces);
            lists = pieces * SP_EPI_HALVES;
            if (int rc = fm3d_scratch(ctx, 4, sizeof(Cand4) * (size_t)lists * nq, (void**)&partial4)) return rc;
            if (SP_N == 256) {
                FM3D_CUDA(ctx, cudaFuncSetAttribute(match_sp_persistent_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.smem_persistent));
                match_sp_persistent_kernel<256><<<G, SP_THREADS, lay.smem_persistent, ctx->stream>>>(ops, ops + ba, nq, q_tiles, nt_tiles, pieces, (dim + 15) / 16, lay, partial4);
            } else {
                FM3D_CUDA(ctx, cudaFuncSetAttribute(match_sp_persistent_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.smem_persistent));
                match_sp_persistent_kernel<128><<<G, SP_THREADS, lay.smem_persistent, ctx->stream>>>(ops, ops + ba, nq, q_tiles, nt_tiles, pieces, (dim + 15) / 16, lay, partial4);
            }
        } else {
            const int splits = pick_splits_waves(q_tiles, nt_tiles, sms, ctx->opt_matcher_splits, &tps);
            lists = splits * SP_EPI_HALVES;
            if (int rc = fm3d_scratch(ctx, 4, sizeof(Cand4) * (size_t)lists * nq, (void**)&partial4)) return rc;
            dim3 grid(q_tiles, splits);
            if (SP_N == 256) {
                FM3D_CUDA(ctx, cudaFuncSetAttribute(match_sp_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.smem));
                match_sp_kernel<256><<<grid, SP_THREADS, lay.smem, ctx->stream>>>(ops, ops + ba, nq, nt_tiles, tps, (dim + 15) / 16, lay, partial4);
            } else {
                FM3D_CUDA(ctx, cudaFuncSetAttribute(match_sp_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.smem));
                match_sp_kernel<128><<<grid, SP_THREADS, lay.smem, ctx->stream>>>(ops, ops + ba, nq, nt_tiles, tps, (dim + 15) / 16, lay, partial4);
            }
        }
        FM3D_LAUNCH_CHECK(ctx);
        sp_refine_kernel<<<(nq + 127) / 128, 128, 0, ctx->stream>>>(q, nq, t, nt, dim, partial4, lists, qnorm, flags + 1, idx, dist,
                                                                    reinterpret_cast<int*>(flags + 2), flagged);
        FM3D_LAUNCH_CHECK(ctx);
        int n_flagged = 0;
        if (int rc = fm3d_d2h(ctx, &n_flagged, flags + 2, sizeof(int))) return rc;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->n_matcher_exact_fallback = n_flagged;
        if (n_flagged > 0) {
            // the filter could not prove these queries: exact brute force for them alone
            auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
            const size_t b_rows = al(sizeof(float) * (size_t)n_flagged * dim), b_i = al(sizeof(int32_t) * 2 * (size_t)n_flagged);
            char* sub = nullptr;
            if (int rc = fm3d_scratch(ctx, 6, b_rows + 2 * b_i, (void**)&sub)) return rc;
            float* qsub = reinterpret_cast<float*>(sub);
            int32_t* sidx = reinterpret_cast<int32_t*>(sub + b_rows);
            float* sdist = reinterpret_cast<float*>(sub + b_rows + b_i);
            sp_gather_rows_kernel<<<n_flagged, 128, 0, ctx->stream>>>(q, dim, flagged, n_flagged, qsub);
            FM3D_LAUNCH_CHECK(ctx);
            if (int rc = knn2_f32_generic(ctx, qsub, n_flagged, t, nt, dim, sidx, sdist)) return rc;
            sp_scatter_kernel<<<(n_flagged + 255) / 256, 256, 0, ctx->stream>>>(flagged, n_flagged, sidx, sdist, idx, dist);
            FM3D_LAUNCH_CHECK(ctx);
        }
        return FM3D_OK;
    }
    return knn2_f32_generic(ctx, q, nq, t, nt, dim, idx, dist);
}

int knn2_ham_dev(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int nbytes,
                 int32_t* idx, float* dist) {
    if (nbytes != 32 && nbytes != 64)
        return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "binary descriptors of %d bytes (supported: 32, 64)", nbytes);
    if (nt >= (1 << 22)) return fm3d_fail(ctx, FM3D_ERR_UNSUPPORTED, "more than 4M train descriptors");
    if ((((uintptr_t)q | (uintptr_t)t) & 15) != 0) return fm3d_fail(ctx, FM3D_ERR_INVALID_ARG, "descriptors must be 16-byte aligned");
    const int sms = ctx->prop.multiProcessorCount;
    const int q_tiles = (nq + 255) / 256, nt_tiles = (nt + 127) / 128;
    int tps = 1;
    const int splits = nt_tiles > 0 ? pick_splits(q_tiles, nt_tiles, 4 * sms, &tps) : 1;
    uint32_t* partial = nullptr;
    if (int rc = fm3d_scratch(ctx, 4, sizeof(uint32_t) * 2 * (size_t)splits * nq, (void**)&partial)) return rc;
    dim3 grid(q_tiles, splits);
    if (nbytes == 32)
        match_ham_kernel<8><<<grid, 128, 0, ctx->stream>>>((const uint32_t*)q, nq, (const uint32_t*)t, nt, tps, partial);
    else
        match_ham_kernel<16><<<grid, 128, 0, ctx->stream>>>((const uint32_t*)q, nq, (const uint32_t*)t, nt, tps, partial);
    FM3D_LAUNCH_CHECK(ctx);
    finalize_ham_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>(partial, splits, nq, nt, idx, dist);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

// host-pointer front end shared by the four host entry points
int match_host(fm3d_ctx* ctx, const void* q, int nq, const void* t, int nt, int row_bytes, int dim_or_bytes,
               bool hamming, bool nndr, double eps, int32_t* idx_or_qidx, int32_t* tidx, float* dist,
               uint8_t* mutual, int* nmatch) {
    if (int rc = fm3d_bind(ctx)) return rc;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t bq = (size_t)nq * row_bytes, bt = (size_t)nt * row_bytes;
    const size_t bi = sizeof(int32_t) * 2 * (size_t)nq, bi_t = sizeof(int32_t) * 2 * (size_t)nt;
    size_t o_q = 0, o_t = o_q + al(bq), o_idx = o_t + al(bt), o_d = o_idx + al(bi);
    size_t o_qi = o_d + al(bi), o_ti = o_qi + al(bi / 2), o_do = o_ti + al(bi / 2), o_n = o_do + al(bi / 2);
    size_t o_iba = o_n + 256, o_dba = o_iba + al(bi_t), o_mu = o_dba + al(bi_t), o_end = o_mu + al((size_t)nq);
    char* d = nullptr;
    if (int rc = fm3d_scratch(ctx, 0, o_end, (void**)&d)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_q, q, bq)) return rc;
    if (int rc = fm3d_h2d(ctx, d + o_t, t, bt)) return rc;
    int32_t* d_idx = (int32_t*)(d + o_idx);
    float* d_dist = (float*)(d + o_d);
    int rc;
    if (nq > 0) {
        rc = hamming ? knn2_ham_dev(ctx, (const uint8_t*)(d + o_q), nq, (const uint8_t*)(d + o_t), nt, dim_or_bytes, d_idx, d_dist)
                     : knn2_f32_dev(ctx, (const float*)(d + o_q), nq, (const float*)(d + o_t), nt, dim_or_bytes, d_idx, d_dist);
        if (rc) return rc;
    }
    if (!nndr) {
        if (int r2 = fm3d_d2h(ctx, idx_or_qidx, d_idx, bi)) return r2;
        if (int r2 = fm3d_d2h(ctx, dist, d_dist, bi)) return r2;
        FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return FM3D_OK;
    }
    nndr_kernel<<<1, 1024, 0, ctx->stream>>>(d_idx, d_dist, nq, eps, (int32_t*)(d + o_qi), (int32_t*)(d + o_ti),
                                             (float*)(d + o_do), (int*)(d + o_n));
    FM3D_LAUNCH_CHECK(ctx);
    if (mutual && nq > 0 && nt > 0) {
        rc = hamming ? knn2_ham_dev(ctx, (const uint8_t*)(d + o_t), nt, (const uint8_t*)(d + o_q), nq, dim_or_bytes,
                                    (int32_t*)(d + o_iba), (float*)(d + o_dba))
                     : knn2_f32_dev(ctx, (const float*)(d + o_t), nt, (const float*)(d + o_q), nq, dim_or_bytes,
                                    (int32_t*)(d + o_iba), (float*)(d + o_dba));
        if (rc) return rc;
        mutual_kernel<<<(nq + 255) / 256, 256, 0, ctx->stream>>>((const int32_t*)(d + o_qi), (const int32_t*)(d + o_ti),
                                                                 (const int*)(d + o_n), (const int32_t*)(d + o_iba),
                                                                 (uint8_t*)(d + o_mu));
        FM3D_LAUNCH_CHECK(ctx);
    }
    int h_n = 0;
    if (int r2 = fm3d_d2h(ctx, &h_n, d + o_n, sizeof(int))) return r2;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (int r2 = fm3d_d2h(ctx, idx_or_qidx, d + o_qi, sizeof(int32_t) * (size_t)h_n)) return r2;
    if (int r2 = fm3d_d2h(ctx, tidx, d + o_ti, sizeof(int32_t) * (size_t)h_n)) return r2;
    if (int r2 = fm3d_d2h(ctx, dist, d + o_do, sizeof(float) * (size_t)h_n)) return r2;
    if (mutual && nq > 0 && nt > 0) if (int r2 = fm3d_d2h(ctx, mutual, d + o_mu, (size_t)h_n)) return r2;
    FM3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *nmatch = h_n;
    return FM3D_OK;
}

}  // namespace

extern "C" {

int fm3d_match_knn2_f32_dev(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                            int32_t* idx, float* dist) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nt >= 0 && dim > 0 && (nq == 0 || (q && idx && dist)) && (nt == 0 || t));
    if (nq == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    return knn2_f32_dev(ctx, q, nq, t, nt, dim, idx, dist);
}

int fm3d_match_knn2_hamming_dev(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt,
                                int nbytes, int32_t* idx, float* dist) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nt >= 0 && nbytes > 0 && (nq == 0 || (q && idx && dist)) && (nt == 0 || t));
    if (nq == 0) return FM3D_OK;
    if (int rc = fm3d_bind(ctx)) return rc;
    return knn2_ham_dev(ctx, q, nq, t, nt, nbytes, idx, dist);
}

int fm3d_nndr_filter_dev(fm3d_ctx* ctx, const int32_t* idx, const float* dist, int nq, double eps,
                         int32_t* qidx, int32_t* tidx, float* dist_out, int* nmatch_dev) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nmatch_dev && (nq == 0 || (idx && dist && qidx && tidx && dist_out)));
    if (int rc = fm3d_bind(ctx)) return rc;
    nndr_kernel<<<1, 1024, 0, ctx->stream>>>(idx, dist, nq, eps, qidx, tidx, dist_out, nmatch_dev);
    FM3D_LAUNCH_CHECK(ctx);
    return FM3D_OK;
}

int fm3d_match_knn2_f32(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                        int32_t* idx, float* dist) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nt >= 0 && dim > 0 && (nq == 0 || (q && idx && dist)) && (nt == 0 || t));
    if (nq == 0) return FM3D_OK;
    return match_host(ctx, q, nq, t, nt, dim * (int)sizeof(float), dim, false, false, 0.0, idx, nullptr, dist, nullptr, nullptr);
}

int fm3d_match_knn2_hamming(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt,
                            int nbytes, int32_t* idx, float* dist) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nt >= 0 && nbytes > 0 && (nq == 0 || (q && idx && dist)) && (nt == 0 || t));
    if (nq == 0) return FM3D_OK;
    return match_host(ctx, q, nq, t, nt, nbytes, nbytes, true, false, 0.0, idx, nullptr, dist, nullptr, nullptr);
}

int fm3d_match_nndr_f32(fm3d_ctx* ctx, const float* q, int nq, const float* t, int nt, int dim,
                        double eps, int32_t* qidx, int32_t* tidx, float* dist, uint8_t* mutual,
                        int* nmatch) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nt >= 0 && dim > 0 && nmatch && (nq == 0 || (q && qidx && tidx && dist)) && (nt == 0 || t));
    *nmatch = 0;
    if (nq == 0) return FM3D_OK;
    return match_host(ctx, q, nq, t, nt, dim * (int)sizeof(float), dim, false, true, eps, qidx, tidx, dist, mutual, nmatch);
}

int fm3d_match_nndr_hamming(fm3d_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt,
                            int nbytes, double eps, int32_t* qidx, int32_t* tidx, float* dist,
                            uint8_t* mutual, int* nmatch) {
    if (!ctx) return FM3D_ERR_INVALID_ARG;
    FM3D_CHECK_ARG(ctx, nq >= 0 && nt >= 0 && nbytes > 0 && nmatch && (nq == 0 || (q && qidx && tidx && dist)) && (nt == 0 || t));
    *nmatch = 0;
    if (nq == 0) return FM3D_OK;
    return match_host(ctx, q, nq, t, nt, nbytes, nbytes, true, true, eps, qidx, tidx, dist, mutual, nmatch);
}

}  // extern "C"
