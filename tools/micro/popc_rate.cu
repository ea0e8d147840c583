// Micro-benchmark: POPC issue rate on B200 (roofline denominator of the Hamming matcher).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o popc_rate popc_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out, int iters, unsigned a) {
    unsigned acc[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { acc[i] = 0; x[i] = threadIdx.x * 2654435761u + i * 40503u; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned p;
            asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(x[i] ^ a));
            acc[i] += p;
            x[i] += acc[i];          // keeps the chain data-dependent without extra popc
        }
    }
    unsigned s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, iters = 20000;
    unsigned* out; cudaMalloc(&out, sizeof(unsigned) * sms * 4 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int threads = 256; threads <= 1024; threads *= 2) {
        for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k<<<sms * 2, threads>>>(out, iters, 0x5a5a5a5au); cudaEventRecord(e1); cudaEventSynchronize(e1); }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)sms * 2 * threads * 8.0 * iters, clk = p.clockRate * 1e3;
        printf("threads/CTA %d x2 CTA/SM: %.3f ms, %.1f popc/clk/SM (at %.0f MHz nominal), %.2f T popc/s (each with 1 xor + 2 add)\n", threads, ms,
               ops / (ms * 1e-3) / clk / sms, clk / 1e6, ops / (ms * 1e-3) / 1e12);
    }
    return 0;
}
