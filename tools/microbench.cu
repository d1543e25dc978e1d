// Micro-benchmarks that calibrate the family-count kernel design on a B200 (run via gpurun):
// shared-memory atomic rates by table size, thread-private counter rates, match.any cost,
// L2 atomic rates and the streaming rate of the kernel's own load pattern.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef unsigned int u32;

__device__ __forceinline__ u32 lcg(u32 &x) { x = x * 1664525u + 1013904223u; return x >> 8; }

// ---- A: shared atomicAdd, pseudo-random cells in [0, C)  (C power of two) ----------------
template <int SKEW>
__global__ void k_atoms(u32 *out, int C, int iters) {
    extern __shared__ u32 h[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u32 x = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    u32 mask = C - 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            u32 c = lcg(x);
            if (SKEW) c = c & (c >> 7) & (c >> 13);   // skewed towards small indices / few hot cells
            atomicAdd(&h[c & mask], 1u);
        }
    }
    __syncthreads();
    u32 s = 0;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s += h[i];
    if (s == 0xdeadbeef) out[0] = s;
}

// ---- C: per-lane private counters, plain LDS / IADD / STS --------------------------------
// layout: warp w owns [C][32] words; lane l touches word cell*32 + l  (bank = lane: conflict-free)
template <int BITS>   // 32: one counter per word; 16: two cells per word
__global__ void k_private(u32 *out, int C, int iters) {
    extern __shared__ u32 h[];
    int words_per_warp = (BITS == 32 ? C : C / 2) * 32;
    for (int i = threadIdx.x; i < words_per_warp * (int)(blockDim.x / 32); i += blockDim.x) h[i] = 0;
    __syncthreads();
    u32 *mine = h + (threadIdx.x >> 5) * words_per_warp + (threadIdx.x & 31);
    u32 x = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    u32 mask = C - 1;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            u32 c = lcg(x) & mask;
            if (BITS == 32) {
                mine[c * 32] += 1u;
            } else {
                mine[(c >> 1) * 32] += 1u << ((c & 1) * 16);
            }
        }
    }
    __syncthreads();
    u32 s = 0;
    for (int i = threadIdx.x; i < words_per_warp * (int)(blockDim.x / 32); i += blockDim.x) s += h[i];
    if (s == 0xdeadbeef) out[0] = s;
}

// ---- D: match.any aggregation then one atomic per distinct cell ----------------------------
__global__ void k_match(u32 *out, int C, int iters) {
    extern __shared__ u32 h[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) h[i] = 0;
    __syncthreads();
    u32 x = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    u32 mask = C - 1;
    int lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            u32 c = lcg(x) & mask;
            u32 peers = __match_any_sync(0xffffffffu, c);
            if (lane == __ffs(peers) - 1) atomicAdd(&h[c], (u32)__popc(peers));
        }
    }
    __syncthreads();
    u32 s = 0;
    for (int i = threadIdx.x; i < C; i += blockDim.x) s += h[i];
    if (s == 0xdeadbeef) out[0] = s;
}

// ---- E: global RED, random cells ----------------------------------------------------------
__global__ void k_red(u32 *tab, u32 mask, int iters) {
    u32 x = blockIdx.x * 7919u + threadIdx.x * 104729u + 1u;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) atomicAdd(&tab[lcg(x) & mask], 1u);
    }
}

// ---- F: streaming read, K+1 columns, 16 B per thread per column ----------------------------
__device__ __forceinline__ uint4 ld_stream_v4(const uint8_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
template <int COLS>
__global__ void k_stream(const uint8_t *data, long long stride, long long nvec, u32 *out) {
    u32 acc = 0;
    long long per = (nvec + gridDim.x - 1) / gridDim.x;
    long long v0 = per * blockIdx.x, v1 = min(nvec, v0 + per);
    for (long long v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
        uint4 w[COLS];
#pragma unroll
        for (int a = 0; a < COLS; ++a) w[a] = ld_stream_v4(data + a * stride + v * 16);
#pragma unroll
        for (int a = 0; a < COLS; ++a) acc += w[a].x ^ w[a].y ^ w[a].z ^ w[a].w;
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 3) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
    u32 *out; CK(cudaMalloc(&out, 1 << 20));
    const double need[] = {6538.0 / 1, 6538.0 / 2, 6538.0 / 3, 6538.0 / 5, 6538.0 / 7};
    printf("rows/s needed at 100%% of 6538 GB/s: k=0 %.0f G, k=1 %.0f G, k=2 %.0f G, k=4 %.0f G, k=6 %.0f G\n", need[0], need[1], need[2], need[3], need[4]);

    const int iters = 2000;
    printf("\n[A] shared atomicAdd (ATOMS), uniform random cells; G ops/s whole GPU\n");
    CK(cudaFuncSetAttribute(k_atoms<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_atoms<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int C : {4, 16, 64, 256, 2048, 8192, 32768}) {
        for (int threads : {256, 1024}) {
            size_t smem = (size_t)C * 4;
            int per_sm = 2048 / threads;
            if (smem * per_sm > 200 * 1024) per_sm = (int)(200 * 1024 / smem);
            if (per_sm < 1) continue;
            int grid = sms * per_sm;
            double ops = (double)grid * threads * iters * 8;
            float ms0 = time_ms([&] { k_atoms<0><<<grid, threads, smem>>>(out, C, iters); });
            float ms1 = time_ms([&] { k_atoms<1><<<grid, threads, smem>>>(out, C, iters); });
            printf("  C=%6d threads=%4d ctas/sm=%d : uniform %8.1f G/s   skewed %8.1f G/s\n", C, threads, per_sm, ops / ms0 * 1e-6, ops / ms1 * 1e-6);
        }
    }

    printf("\n[C] per-lane private counters (LDS+IADD+STS), uniform random cells\n");
    CK(cudaFuncSetAttribute(k_private<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    CK(cudaFuncSetAttribute(k_private<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int C : {4, 16, 64, 128, 256, 512}) {
        for (int threads : {128, 256, 512, 1024}) {
            int warps = threads / 32;
            size_t s32 = (size_t)C * 128 * warps, s16 = s32 / 2;
            auto run = [&](size_t smem, int bits) {
                int per_sm = 2048 / threads;
                if (smem * per_sm > 200 * 1024) per_sm = (int)(200 * 1024 / smem);
                if (per_sm < 1) { printf("      -    "); return; }
                int grid = sms * per_sm;
                double ops = (double)grid * threads * iters * 8;
                float ms = bits == 32 ? time_ms([&] { k_private<32><<<grid, threads, smem>>>(out, C, iters); })
                                      : time_ms([&] { k_private<16><<<grid, threads, smem>>>(out, C, iters); });
                printf(" %8.1f G/s (x%d)", ops / ms * 1e-6, per_sm);
            };
            printf("  C=%4d threads=%4d : u32", C, threads);
            run(s32, 32);
            printf("   u16");
            run(s16, 16);
            printf("\n");
        }
    }

    printf("\n[D] match.any + one atomic per distinct cell\n");
    for (int C : {4, 16, 256, 8192}) {
        int threads = 256, grid = sms * 8;
        double ops = (double)grid * threads * iters * 8;
        float ms = time_ms([&] { k_match<<<grid, threads, C * 4>>>(out, C, iters); });
        printf("  C=%6d : %8.1f G/s\n", C, ops / ms * 1e-6);
    }

    printf("\n[E] global RED (L2 atomics), uniform random cells\n");
    u32 *tab; CK(cudaMalloc(&tab, (size_t)1 << 28));
    CK(cudaMemset(tab, 0, (size_t)1 << 28));
    for (u32 cells : {1u << 10, 1u << 16, 1u << 20, 1u << 26}) {
        int threads = 256, grid = sms * 8, it = 200;
        double ops = (double)grid * threads * it * 8;
        float ms = time_ms([&] { k_red<<<grid, threads>>>(tab, cells - 1, it); });
        printf("  cells=%9u : %8.1f G/s\n", cells, ops / ms * 1e-6);
    }

    printf("\n[F] streaming read of COLS columns, 16 B/thread/column (GB/s)\n");
    long long N = 1ll << 28;   // 256 Mi rows per column, 7 columns = 1.75 GiB
    uint8_t *data; CK(cudaMalloc(&data, (size_t)N * 7));
    CK(cudaMemset(data, 1, (size_t)N * 7));
    long long nvec = N / 16;
    for (int threads : {256, 512, 1024}) {
        for (int per_sm : {1, 2, 4, 8}) {
            if (threads * per_sm > 2048) continue;
            int grid = sms * per_sm;
            float m1 = time_ms([&] { k_stream<1><<<grid, threads>>>(data, N, nvec, out); });
            float m3 = time_ms([&] { k_stream<3><<<grid, threads>>>(data, N, nvec, out); });
            float m5 = time_ms([&] { k_stream<5><<<grid, threads>>>(data, N, nvec, out); });
            float m7 = time_ms([&] { k_stream<7><<<grid, threads>>>(data, N, nvec, out); });
            printf("  threads=%4d ctas/sm=%d : cols=1 %7.0f  cols=3 %7.0f  cols=5 %7.0f  cols=7 %7.0f\n", threads, per_sm,
                   N * 1.0 / m1 * 1e-6, N * 3.0 / m3 * 1e-6, N * 5.0 / m5 * 1e-6, N * 7.0 / m7 * 1e-6);
        }
    }
    printf("done\n");
    return 0;
}
