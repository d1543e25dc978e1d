// Streaming k+1 uint8 columns: direct 128-bit loads vs TMA bulk copies (cp.async.bulk, mbarrier
// completion) staged through a shared-memory ring and read back with LDS.128.  Answers whether
// TMA staging would help the family-count kernel, whose shared-memory data pipe is the limiter.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef unsigned int u32;
typedef unsigned long long u64;

__device__ __forceinline__ uint4 ld_stream_v4(const uint8_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile("{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

template <int COLS>
__global__ void k_stream_ldg(const uint8_t *data, long long stride, long long nvec, u32 *out) {
    u32 acc = 0;
    long long per = ((nvec + gridDim.x - 1) / gridDim.x + 7) & ~7ll;
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

// One tile = THREADS * 16 bytes per column (each thread consumes one 16-byte vector per column).
template <int COLS, int THREADS, int STAGES>
__global__ void __launch_bounds__(THREADS) k_stream_tma(const uint8_t *data, long long stride, long long nvec, u32 *out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ u64 full[STAGES];
    constexpr int TILE = THREADS * 16;
    long long per = ((nvec + gridDim.x - 1) / gridDim.x + 7) & ~7ll;
    long long v0 = per * blockIdx.x, v1 = min(nvec, v0 + per);
    long long ntiles = (v1 - v0 + THREADS - 1) / THREADS;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long t) {
        int s = (int)(t % STAGES);
        long long vb = v0 + t * THREADS;
        u32 bytes = (u32)(min((long long)THREADS, v1 - vb) * 16);
        mbar_expect_tx(&full[s], bytes * COLS);
        for (int a = 0; a < COLS; ++a) bulk_g2s(smem + ((size_t)s * COLS + a) * TILE, data + a * stride + vb * 16, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (long long t = 0; t < STAGES - 1 && t < ntiles; ++t) issue(t);
    u32 acc = 0;
    for (long long t = 0; t < ntiles; ++t) {
        int s = (int)(t % STAGES);
        if (threadIdx.x == 0 && t + STAGES - 1 < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(t + STAGES - 1);
        }
        mbar_wait(&full[s], (u32)((t / STAGES) & 1));
        long long v = v0 + t * THREADS + threadIdx.x;
        if (v < v1) {
#pragma unroll
            for (int a = 0; a < COLS; ++a) {
                uint4 w = *reinterpret_cast<const uint4 *>(smem + ((size_t)s * COLS + a) * TILE + threadIdx.x * 16);
                acc += w.x ^ w.y ^ w.z ^ w.w;
            }
        }
        __syncthreads();   // everyone is done with stage s before it is refilled
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}

template <typename F> float time_ms(F f) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms; }
    CK(cudaGetLastError());
    return best;
}

template <int COLS> void run(const uint8_t *data, long long N, u32 *out, int sms) {
    long long nvec = N / 16;
    constexpr int THREADS = 256, STAGES = 4;
    size_t smem = (size_t)STAGES * COLS * THREADS * 16;
    CK(cudaFuncSetAttribute(k_stream_tma<COLS, THREADS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int per_sm : {2, 4}) {
        if (smem * per_sm > 200 * 1024) continue;
        int grid = sms * per_sm;
        float a = time_ms([&] { k_stream_ldg<COLS><<<grid, THREADS>>>(data, N, nvec, out); });
        float b = time_ms([&] { k_stream_tma<COLS, THREADS, STAGES><<<grid, THREADS, smem>>>(data, N, nvec, out); });
        printf("  cols=%d ctas/sm=%d : LDG.128 %7.0f GB/s   TMA bulk (%d stages x %zu KB) + LDS.128 %7.0f GB/s\n", COLS, per_sm,
               N * (double)COLS / a * 1e-6, STAGES, smem / STAGES / 1024, N * (double)COLS / b * 1e-6);
    }
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    u32 *out; CK(cudaMalloc(&out, 4096));
    long long N = 1ll << 28;
    uint8_t *data; CK(cudaMalloc(&data, (size_t)N * 7));
    CK(cudaMemset(data, 1, (size_t)N * 7));
    printf("streaming COLS uint8 columns of %lld rows, 256 threads per CTA\n", N);
    run<1>(data, N, out, sms);
    run<3>(data, N, out, sms);
    run<5>(data, N, out, sms);
    run<7>(data, N, out, sms);
    return 0;
}
