// Micro-benchmark: FP32 FMA issue rate on B200, scalar FFMA vs packed FFMA2 (fma.rn.f32x2).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
        }
    } else {
        float2 av = make_float2(a, a), bv = make_float2(b, b);
        u64 A = *reinterpret_cast<u64*>(&av), B = *reinterpret_cast<u64*>(&bv);
        u64 p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 t = make_float2(acc[2 * i], acc[2 * i + 1]); p[i] = *reinterpret_cast<u64*>(&t); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = ffma2(p[i], A, B);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 t = *reinterpret_cast<float2*>(&p[i]); acc[2 * i] = t.x; acc[2 * i + 1] = t.y; }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, iters = 20000;
    float* out; cudaMalloc(&out, sizeof(float) * sms * 4 * 512);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads = 128; threads <= 512; threads *= 2)
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<sms * 2, threads>>>(out, iters, 1.0001f, 1e-7f); else k<1><<<sms * 2, threads>>>(out, iters, 1.0001f, 1e-7f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)sms * 2 * threads * 16.0 * iters;
        double clk = p.clockRate * 1e3;   // Hz
        printf("threads/CTA %d x2 CTA/SM, %s: %.3f ms, %.1f FMA/clk/SM (at %.0f MHz nominal), %.2f TFLOP/s\n", threads, mode ? "FFMA2" : "FFMA ",
               ms, fma / (ms * 1e-3) / clk / sms, clk / 1e6, 2 * fma / (ms * 1e-3) / 1e12);
    }
    return 0;
}
