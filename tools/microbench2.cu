// Shared-memory atomic wavefront model on B200: fixed address patterns, ATOMS.POPC.INC (+1) vs
// ATOMS.ADD (+2), to learn what the atomic unit treats as a bank conflict.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef unsigned int u32;
__device__ __forceinline__ u32 lcg(u32 &x) { x = x * 1664525u + 1013904223u; return x >> 8; }

// PATTERN: word index as a function of lane and a per-op pseudo-random value
template <int PATTERN, int INC>
__global__ void k_pat(u32 *out, int iters) {
    extern __shared__ u32 h[];
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u32 lane = threadIdx.x & 31;
    u32 x = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            u32 c = lcg(x);
            u32 w;
            if (PATTERN == 0) w = lane;                          // 32 distinct banks, fixed
            else if (PATTERN == 1) w = lane * 2;                 // stride 8 B
            else if (PATTERN == 2) w = 0;                        // one address
            else if (PATTERN == 3) w = lane & 15;                // pairs share an address
            else if (PATTERN == 4) w = lane * 32;                // one bank, 32 addresses
            else if (PATTERN == 5) w = (c & 63) * 32 + lane;     // random cell, 32 lane replicas (bank = lane)
            else if (PATTERN == 6) w = c & 2047;                 // random cell, no replicas
            else if (PATTERN == 7) w = (c & 255) * 32 + lane;    // 256 cells x 32 replicas
            else if (PATTERN == 8) w = (c & 63) * 64 + lane * 2; // replicas on 8-byte stride
            else if (PATTERN == 9) w = (c & 511) * 16 + (lane & 15);   // 16 replicas
            else w = (c & 63);                                   // 64 cells, no replicas
            if (PATTERN < 5 && PATTERN != 2) w += (c & 1) * 0;   // keep c live
            atomicAdd(&h[w], (u32)INC);
        }
    }
    __syncthreads();
    u32 s = 0;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) s += h[i];
    if (s == 0xdeadbeef) out[0] = s;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

template <int P> void run(const char *name, u32 *out, int sms) {
    const int iters = 2000, threads = 256, per_sm = 3, grid = sms * per_sm;
    size_t smem = 65536;
    CK(cudaFuncSetAttribute(k_pat<P, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_pat<P, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double ops = (double)grid * threads * iters * 8;
    float m1 = time_ms([&] { k_pat<P, 1><<<grid, threads, smem>>>(out, iters); });
    float m2 = time_ms([&] { k_pat<P, 2><<<grid, threads, smem>>>(out, iters); });
    double clk = 1.965e9;
    printf("  %-52s POPC.INC %7.1f G/s (%.2f clk/warp-op/SM)   ADD %7.1f G/s (%.2f)\n", name, ops / m1 * 1e-6,
           32.0 / (ops / (m1 * 1e-3) / sms / clk), ops / m2 * 1e-6, 32.0 / (ops / (m2 * 1e-3) / sms / clk));
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    u32 *out; CK(cudaMalloc(&out, 4096));
    printf("shared atomics, 768 threads/SM, clk/warp-op assumes 1.965 GHz\n");
    run<0>("lane (32 banks, distinct words)", out, sms);
    run<1>("lane*2 (8-byte stride)", out, sms);
    run<2>("one address", out, sms);
    run<3>("lane&15 (pairs share a word)", out, sms);
    run<4>("lane*32 (one bank, 32 words)", out, sms);
    run<5>("rand64*32 + lane (32 replicas, bank = lane)", out, sms);
    run<7>("rand256*32 + lane (32 replicas)", out, sms);
    run<8>("rand64*64 + lane*2 (replicas, 8-byte stride)", out, sms);
    run<9>("rand512*16 + lane&15 (16 replicas)", out, sms);
    run<6>("rand2048 (no replicas)", out, sms);
    run<10>("rand64 (no replicas)", out, sms);
    return 0;
}
