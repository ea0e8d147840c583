// Micro-benchmark: sustained FFMA2 (fma.rn.f32x2) rate on B200 as a function of its register operands.
// The peak figure of ffma2_rate.cu re-uses two of the three 64-bit sources; the pixel loop of the normal
// search (fm3d_normals_fast.cu) mostly has three distinct sources, one of them a 32-bit broadcast scalar.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_operands ffma2_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ u64 ffma2_bc(float a, u64 b, u64 c) { return ffma2(mk2(a, a), b, c); }
__device__ __forceinline__ u64 ffma2_bc2(float a, u64 b, float c) { return ffma2(mk2(a, a), b, mk2(c, c)); }

// MODE 0: d = d*A + B (two sources shared by all chains: the peak benchmark)
// MODE 1: d = d*q_i + B        (two distinct 64-bit sources per instruction)
// MODE 2: d = d*q_i + r_i      (three distinct 64-bit sources)
// MODE 3: d = bc(s_i)*d + r_i  (32-bit broadcast scalar + two 64-bit sources: the Horner / homography form)
// MODE 4: d = bc(s_i)*d + bc(t_i) (two broadcast scalars)
// MODE 5: scalar FFMA with three distinct sources: d = d*q_i + r_i on 16 chains
template <int MODE>
__global__ void k(float* out, int iters, float a, float b, const float* __restrict__ src) {
    // sources come from memory (per-thread addresses): ptxas cannot prove them warp-uniform, so they live in
    // ordinary registers as the pass constants of the normal search do (they are read from shared memory)
    u64 p[8], q[8], r[8];
    float s[8], t[8];
    const float* g = src + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p[i] = mk2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
        q[i] = mk2(a + g[i * 512], a - g[i * 512]);
        r[i] = mk2(b + g[(i + 8) * 512], b - g[(i + 8) * 512]);
        s[i] = a + g[(i + 16) * 512]; t[i] = b + g[(i + 24) * 512];
    }
    float f[16], fq[16], fr[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { f[i] = threadIdx.x * 1e-3f + i; fq[i] = a + g[i * 512]; fr[i] = b + g[(i + 16) * 512]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) p[i] = ffma2(p[i], q[0], r[0]);
            if (MODE == 1) p[i] = ffma2(p[i], q[i], r[0]);
            if (MODE == 2) p[i] = ffma2(p[i], q[i], r[i]);
            if (MODE == 3) p[i] = ffma2_bc(s[i], p[i], r[i]);
            if (MODE == 4) p[i] = ffma2_bc2(s[i], p[i], t[i]);
        }
        if (MODE == 5) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fq[i]), "f"(fr[i]));
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)); }
#pragma unroll
    for (int i = 0; i < 16; i++) acc += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, int sms, int clock_khz, float* out) {
    const int iters = 20000, threads = 512;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<sms, threads>>>(out, iters, 1.0001f, 1e-7f, out + sms * 512);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)sms * threads * 16.0 * iters, clk = clock_khz * 1e3;
    printf("%-58s %.3f ms  %.1f FMA/clk/SM  %.2f TFLOP/s  (%.2f clk per warp instruction per SMSP)\n", name, ms, fma / (ms * 1e-3) / clk / sms,
           2 * fma / (ms * 1e-3) / 1e12, (MODE == 5 ? 32.0 : 64.0) / (fma / (ms * 1e-3) / clk / sms / 4));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float* out; cudaMalloc(&out, sizeof(float) * (p.multiProcessorCount * 512 + 64 * 512)); cudaMemset(out, 0, sizeof(float) * (p.multiProcessorCount * 512 + 64 * 512));
    printf("# %s, %d SMs, %d MHz, 512 threads x 1 CTA/SM (16 warps/SM, as the normal-search kernel)\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
    run<0>("FFMA2 d = d*q_0 + r_0    (2 of 3 sources shared)", p.multiProcessorCount, p.clockRate, out);
    run<1>("FFMA2 d = d*q_i + r_0    (2 distinct + 1 shared)", p.multiProcessorCount, p.clockRate, out);
    run<2>("FFMA2 d = d*q_i + r_i    (3 distinct 64-bit sources)", p.multiProcessorCount, p.clockRate, out);
    run<3>("FFMA2 d = bc(s_i)*d + r_i (broadcast scalar + 2 x 64-bit)", p.multiProcessorCount, p.clockRate, out);
    run<4>("FFMA2 d = bc(s_i)*d + bc(t_i) (2 broadcast scalars)", p.multiProcessorCount, p.clockRate, out);
    run<5>("FFMA  d = d*q_i + r_i    (scalar, 3 distinct sources)", p.multiProcessorCount, p.clockRate, out);
    return 0;
}
