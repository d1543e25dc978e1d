// libbicgpu.so — host side of the C ABI declared in include/bicgpu.h.
//
// Drop-in for the reference's per-DAG Rscript child (bnlearn.py:46-61 ->
// bnlearn_score.R:7-40).  The context owns: the padded column-major uint8 dataset in HBM, the
// family-score cache (open-addressing table + registry of keys / log-likelihoods), per-batch
// workspace and a pinned header the kernels report through.  There is no CPU fallback: every
// score comes from the kernels in count_kernels.cuh.
#include "../../include/bicgpu.h"

#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "cache_kernels.cuh"
#include "common.cuh"
#include "count_kernels.cuh"

using namespace bic;

namespace {

thread_local std::string g_create_error;   // bic_last_error(NULL): last bic_create failure of the calling thread

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 2 + 256;   // growth must be rare: cudaFree + cudaMalloc take milliseconds
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ---- NCCL through dlopen: single-GPU users never need the library -------------------------
typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_uid, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    std::string err;
    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
        GetUniqueId = (int (*)(nccl_uid *))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (int (*)(nccl_comm *, int, nccl_uid, int))dlsym(lib, "ncclCommInitRank");
        AllReduce = (int (*)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t))dlsym(lib, "ncclAllReduce");
        AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm, cudaStream_t))dlsym(lib, "ncclAllGather");
        CommDestroy = (int (*)(nccl_comm))dlsym(lib, "ncclCommDestroy");
        GetErrorString = (const char *(*)(int))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !AllReduce || !AllGather || !CommDestroy) {
            err = "libnccl lacks a required symbol";
            lib = nullptr;
            return false;
        }
        return true;
    }
};
NcclApi g_nccl;
std::mutex g_nccl_mu;
constexpr int NCCL_UINT8 = 1, NCCL_UINT32 = 3, NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MIN = 3;   // nccl.h: ncclDataType_t / ncclRedOp_t

}  // namespace

struct bic_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::mutex mu;
    std::string err;
    int sm_count = 148;
    size_t attr_smem[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // opt-in dynamic shared memory already set per k_count instance

    // dataset
    uint8_t *data = nullptr;
    uint8_t *data2 = nullptr;    // 2-bit packed shadow copy (columns with <= 4 states), stride2 = stride / 4
    bool all_packed = false;     // every column has one: all families stream the packed copy
    long long stride2 = 0;
    long long N = 0, stride = 0, N_total = 0;
    int n = 0, W64 = 0, Wk = 0;
    int *d_card = nullptr;
    std::vector<int> card;

    // family-score cache
    u32 *table = nullptr;
    u64 table_cap = 0;
    u64 *regkeys = nullptr;
    double *reg_ll = nullptr, *reg_np = nullptr;
    long long reg_cap = 0, reg_count = 0;
    long long lookups = 0, misses = 0;
    int cache_mode = 0;      // what reg_ll holds: 0 log-likelihood (+ reg_np), 1 BDeu terms for `iss`, 2 K2 terms
    double iss = 1.0;

    // per-sub-batch workspace
    DevBuf keybuf, inst, flag, rank, bsum32, bsum64, cells_arr, class_jobs, need, table_off, done, arena;
    DevBuf dag_bad, in_stage, in_stage2, in_stage3, in_stage4, out_stage, tmp_ll, donor, donor_best, derived_list, derived_sorted, owner, xoff, fp_buf;
    // Tuning knobs.  Defaults are the values swept on B200; the BIC_*
    // environment variables exist for those sweeps and for tests, not as a supported interface.
    struct Tuning {
        bool derive = true;                    // BIC_NO_DERIVE=1: count every family from the rows
        bool pack2 = true;                     // BIC_NO_PACK2=1: no 2-bit shadow copy
        long long pack2_min_rows = 1ll << 20;  // BIC_PACK2_MIN_ROWS
        long long derive_min_rows = 1ll << 20;
        long long l2_window = 32ll << 20;      // BIC_L2_WINDOW_MB: dataset bytes of one row slice kept L2-resident
        long long l2_window_max = 256ll << 20; // BIC_L2_WINDOW_MAX_MB: window when all families of a class are resident at once
        u32 class0_words = CLASS0_WORDS;       // BIC_CLASS0_WORDS: shared-memory words of a class-0 CTA (uint8 path)
        u32 class0_words_packed = CLASS0_WORDS_PACKED;   // the same when every column streams from the 2-bit packed copy
        int class0_threads = 256;              // BIC_CLASS0_THREADS: 256, 512 or 1024
        bool class0_wide = true;               // BIC_CLASS0_WIDE=0: never the 512-thread x 96 KB class-0 shape
        bool class0_explicit = false;          // BIC_CLASS0_WORDS / BIC_CLASS0_THREADS given: no automatic wide shape (class0_shape)
        int range_passes = 8;                  // BIC_RANGE_PASSES: class-3 tables of up to this many shared-memory sub-ranges
                                               //   are counted in passes (0: always straight into HBM with L2 atomics)
        int class1_threads = 512;              // BIC_CLASS1_THREADS: 256, 512 or 1024 (48 KB tables)
        int class2_threads = 1024;             // BIC_CLASS2_THREADS: 512 or 1024 (classes 2 and 3-in-passes: one CTA per SM)
        int p2_vec = 4;                        // BIC_P2_VEC: 32-bit words of a packed column per thread-iteration (4, 2 or 1)
        int cluster = 0;                       // BIC_CLUSTER=1: class 3 in one pass over a thread-block cluster (measured 3x slower than sub-range passes)
        int cluster_size = 0;                  // BIC_CLUSTER_SIZE: force 2, 4 or 8 CTAs per cluster (0: smallest that holds the table)
        int cluster_threads = 1024;            // BIC_CLUSTER_THREADS: 512 or 1024
        bool park_cells = false;               // BIC_PARK_CELLS=1: class-3 passes read the cell index of every row from scratch that k_cells
                                               //   fills once, instead of recomputing it per pass (measured slower: 0.57 vs 0.38 ms)
        long long cells_max_mb = 4096;         // BIC_CELLS_MAX_MB: scratch limit; above it the passes recompute
        bool sort_jobs = false;                // BIC_SORT_JOBS=1: count jobs of a class ordered by their number of parents (measured 2 % slower:
                                               //   co-resident CTAs of different shapes load the ALU and the shared-memory pipe more evenly)
        bool park_meta = true;                 // BIC_NO_META=1: thread 0 of every count CTA decodes its family key (round-1 behaviour)
        int u8_two = 2;                        // BIC_U8_TWO: uint8 path, families of <= 4 columns: row groups in flight per thread
                                               //   (0 / 1: one, 2: two, 3: four / three / two for k = 0 / 1 / >= 2)
        int p2_two = 0;                        // BIC_P2_TWO=1: packed path, families of <= 3 columns keep two 64-row groups in flight
                                               //   (experiment; pigs-shaped class-0 launch 1.033 -> 1.087 ms, 3 runs each: off)
        bool swizzle = true;                   // BIC_SWIZZLE=0: un-replicated shared-memory tables in plain cell order
        u32 tier0 = 768, tier1 = 3072;         // BIC_TIER0 / BIC_TIER1: cell limits of the first launch over the class-0 / class-1 list (0: one launch).
                                               //   Alarm-shaped step 54.45 ms untiered; 53.05 (768, 0); 53.58 (0, 3072); 52.23 (768, 3072); 52.37 (1024, 3072);
                                               //   52.71 (1536, 3072)
        bool list3 = true;                     // BIC_NO_LIST3=1: class-3 grid of njobs x (most passes) items per slice, surplus items empty
        bool c3_u16 = false;                   // BIC_C3_U16=1: class-3 sub-ranges with 16-bit counters (half the passes; count_rows_r16)
        bool topsplit = true;                  // BIC_TOPSPLIT=0: class-3 sub-ranges always by cell index, every pass computes the full index of every row
        bool u8_narrow = false;                // BIC_U8_NARROW=1: uint8 path of classes 0 / 1 loads 8 bytes per thread per column (experiment)
        bool tma = false;                      // BIC_TMA=1: uint8 path of classes 0 / 1 stages its rows with TMA bulk copies (experiment)
        bool push = true;                      // BIC_NO_PUSH=1: row-sharded runs all-reduce the count tables with NCCL instead of the
                                               //   fused reduce-scatter over peer memory
        long long xchg_mb = 64;                // BIC_XCHG_MB: exchange-buffer slot per source rank (MB)
        bool push_world1 = false;              // BIC_PUSH_WORLD1=1 (tests): a one-rank communicator also takes the exchange-buffer path
        bool fast_small = true;                // BIC_NO_FAST_SMALL=1: small warm batches take the general pipeline too
        bool slice_model = true;               // BIC_SLICE_MODEL=0: always cut the rows into L2 windows (round-1 versions a-h)
        void from_env() {
            if (const char *e = getenv("BIC_NO_DERIVE")) derive = atoi(e) == 0;
            if (const char *e = getenv("BIC_NO_PACK2")) pack2 = atoi(e) == 0;
            if (const char *e = getenv("BIC_PACK2_MIN_ROWS")) pack2_min_rows = atoll(e);
            if (const char *e = getenv("BIC_L2_WINDOW_MB")) { long long mb = atoll(e); if (mb > 0) l2_window = mb << 20; }
            if (const char *e = getenv("BIC_L2_WINDOW_MAX_MB")) { long long mb = atoll(e); if (mb > 0) l2_window_max = mb << 20; }
            if (const char *e = getenv("BIC_CLASS0_WORDS")) { int w = atoi(e); if (w >= (int)CLASS0_CELLS && w <= 49152) { class0_words = (u32)w; class0_words_packed = (u32)w; class0_explicit = true; } }
            if (const char *e = getenv("BIC_RANGE_PASSES")) { int v = atoi(e); if (v >= 0 && v <= 64) range_passes = v; }
            if (const char *e = getenv("BIC_CLASS1_THREADS")) { int t = atoi(e); if (t == 256 || t == 512 || t == 1024) class1_threads = t; }
            if (const char *e = getenv("BIC_CLASS2_THREADS")) { int t = atoi(e); if (t == 512 || t == 1024) class2_threads = t; }
            if (const char *e = getenv("BIC_P2_VEC")) { int v = atoi(e); if (v == 1 || v == 2 || v == 4) p2_vec = v; }
            if (const char *e = getenv("BIC_CLUSTER")) cluster = atoi(e) != 0;
            if (const char *e = getenv("BIC_CLUSTER_SIZE")) { int v = atoi(e); if (v == 0 || v == 2 || v == 4 || v == 8) cluster_size = v; }
            if (const char *e = getenv("BIC_CLUSTER_THREADS")) { int v = atoi(e); if (v == 512 || v == 1024) cluster_threads = v; }
            if (const char *e = getenv("BIC_TMA")) tma = atoi(e) != 0;
            if (const char *e = getenv("BIC_U8_NARROW")) u8_narrow = atoi(e) != 0;
            if (const char *e = getenv("BIC_TOPSPLIT")) topsplit = atoi(e) != 0;
            if (const char *e = getenv("BIC_C3_U16")) c3_u16 = atoi(e) != 0;
            if (const char *e = getenv("BIC_NO_LIST3")) list3 = atoi(e) == 0;
            if (const char *e = getenv("BIC_TIER0")) { int v = atoi(e); if (v >= 0 && v < (int)CLASS0_CELLS) tier0 = (u32)v; }
            if (const char *e = getenv("BIC_TIER1")) { int v = atoi(e); if (v >= 0 && v <= 3072) tier1 = (u32)v; }
            if (const char *e = getenv("BIC_SWIZZLE")) swizzle = atoi(e) != 0;
            if (const char *e = getenv("BIC_P2_TWO")) p2_two = atoi(e) != 0;
            if (const char *e = getenv("BIC_U8_TWO")) { int v = atoi(e); u8_two = v <= 1 ? 0 : v >= 3 ? 3 : 2; }
            if (const char *e = getenv("BIC_NO_META")) park_meta = atoi(e) == 0;
            if (const char *e = getenv("BIC_SORT_JOBS")) sort_jobs = atoi(e) != 0;
            if (const char *e = getenv("BIC_PARK_CELLS")) park_cells = atoi(e) != 0;
            if (const char *e = getenv("BIC_CELLS_MAX_MB")) { long long v = atoll(e); if (v >= 0) cells_max_mb = v; }
            if (const char *e = getenv("BIC_NO_PUSH")) push = atoi(e) == 0;
            if (const char *e = getenv("BIC_PUSH_WORLD1")) push_world1 = atoi(e) != 0;
            if (const char *e = getenv("BIC_XCHG_MB")) { long long v = atoll(e); if (v > 0 && v <= 4096) xchg_mb = v; }
            if (const char *e = getenv("BIC_NO_FAST_SMALL")) fast_small = atoi(e) == 0;
            if (const char *e = getenv("BIC_SLICE_MODEL")) slice_model = atoi(e) != 0;
            if (const char *e = getenv("BIC_CLASS0_WIDE")) class0_wide = atoi(e) != 0;
            if (const char *e = getenv("BIC_CLASS0_THREADS")) { int t = atoi(e); if (t == 256 || t == 512 || t == 1024) { class0_threads = t; class0_explicit = true; } }
        }
    } tune;
    Header *d_hdr = nullptr, *h_hdr = nullptr;

    // profiling
    bool prof_on = false;
    bic_profile_t prof = {};
    struct EvPair { cudaEvent_t a, b; int cls; };
    std::vector<EvPair> ev_pool, ev_used;

    // row sharding
    nccl_comm comm = nullptr;
    int rank_id = 0, world = 1;
    int comm_mode = 0;       // BIC_SHARD_ROWS or BIC_SHARD_FAMILIES
    bool ntotal_dirty = true;
    // fused reduce-scatter of row-sharded count tables: every rank owns `world` slots of xcap cells;
    // slot r of rank o receives rank r's partial tables of the families o owns (peer stores over NVLink)
    u32 *xchg = nullptr;                 // this rank's exchange buffer (cudaMalloc, IPC-exported)
    u64 xcap = 0;                        // cells per slot
    std::vector<void *> peer_open;       // peers' buffers opened with cudaIpcOpenMemHandle (to close)
    u32 **d_peer = nullptr;              // device array [world] of exchange-buffer pointers (own buffer at [rank])
    int xchg_state = 0;                  // 0 not tried, 1 ready, -1 unavailable (fall back to ncclAllReduce of the tables)
    int *d_barrier = nullptr;            // 4 bytes all-reduced as the "all pushes have landed" barrier
    DevBuf terms, gkeys, gbad, cellbuf, meta, jobs_sorted, items3;  // staged family terms (one all-reduce), all-gathered keys / reject flags, parked cell indices (class 3)
    cudaEvent_t ev_wait = nullptr;       // bic_wait_stream
    bool fast_ok = true;     // small warm batches: try k_score_small first (off after a miss, on again after an all-hit call)
};

namespace {

int fail(bic_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(c, e_ == cudaErrorMemoryAllocation ? BIC_ERR_OOM : BIC_ERR_CUDA,           \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    } while (0)

#define TRY(call)                         \
    do {                                  \
        int rc_ = (call);                 \
        if (rc_ != BIC_OK) return rc_;    \
    } while (0)

inline u64 mix_host(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

inline unsigned nblk(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

#define LAUNCH(c) (++(c)->prof.kernel_launches)

// Exclusive scan of `in[0..n)` into `out`, total written to *total (device).
template <typename TI, typename TO>
int scan_excl(bic_ctx *c, const TI *in, long long n, TO *out, TO *total, DevBuf &bsum) {
    int nb = (int)((n + SCAN_CHUNK - 1) / SCAN_CHUNK);
    if (nb < 1) nb = 1;
    CU(bsum.ensure((size_t)nb * sizeof(TO)));
    k_scan_partial<TI, TO><<<nb, SCAN_THREADS, 0, c->stream>>>(in, n, bsum.as<TO>()); LAUNCH(c);
    k_scan_bsums<TO><<<1, 1024, 0, c->stream>>>(bsum.as<TO>(), nb, total); LAUNCH(c);
    k_scan_apply<TI, TO><<<nb, SCAN_THREADS, 0, c->stream>>>(in, n, bsum.as<TO>(), out); LAUNCH(c);
    CU(cudaGetLastError());
    return BIC_OK;
}

void cache_free(bic_ctx *c) {
    if (c->table) cudaFree(c->table);
    if (c->regkeys) cudaFree(c->regkeys);
    if (c->reg_ll) cudaFree(c->reg_ll);
    if (c->reg_np) cudaFree(c->reg_np);
    c->table = nullptr; c->regkeys = nullptr; c->reg_ll = nullptr; c->reg_np = nullptr;
    c->table_cap = 0; c->reg_cap = 0; c->reg_count = 0;
}

// Make room for `extra` more families (worst case: every instance of a sub-batch is new).
// Load factor stays <= 0.5; growth re-inserts the registry ids into a fresh table.
int cache_ensure(bic_ctx *c, long long extra) {
    long long want = c->reg_count + extra;
    if (c->table && want <= c->reg_cap) return BIC_OK;
    // cudaMalloc / cudaFree take anywhere from 1 ms to hundreds of ms here (measured), so growth must
    // be rare: start at 2^20 families (40 MB) and leave 2x headroom over what is needed now
    want = std::max<long long>(2 * want, 1ll << 20);
    u64 cap = 1ull << 16;
    while ((long long)(cap / 2) < want) cap <<= 1;
    if (cap > (1ull << 30)) {
        cap = 1ull << 30;
        if ((long long)(cap / 2) < c->reg_count + extra)
            return fail(c, BIC_ERR_OOM, "family cache would exceed 2^30 slots; call bic_cache_clear()");
    }
    long long rcap = (long long)(cap / 2);
    u32 *ntable = nullptr; u64 *nkeys = nullptr; double *nll = nullptr, *nnp = nullptr;
    if (cudaMalloc(&ntable, cap * sizeof(u32)) != cudaSuccess || cudaMalloc(&nkeys, (size_t)rcap * c->Wk * sizeof(u64)) != cudaSuccess ||
        cudaMalloc(&nll, (size_t)rcap * sizeof(double)) != cudaSuccess || cudaMalloc(&nnp, (size_t)rcap * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        if (ntable) cudaFree(ntable);
        if (nkeys) cudaFree(nkeys);
        if (nll) cudaFree(nll);
        if (nnp) cudaFree(nnp);
        return fail(c, BIC_ERR_OOM, "device allocation for the family cache failed");
    }
    // fill the new allocation; on any failure it is released and the old cache stays in place
    cudaError_t e = cudaMemsetAsync(ntable, 0, cap * sizeof(u32), c->stream);
    if (e == cudaSuccess && c->reg_count) {
        e = cudaMemcpyAsync(nkeys, c->regkeys, (size_t)c->reg_count * c->Wk * sizeof(u64), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nll, c->reg_ll, (size_t)c->reg_count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nnp, c->reg_np, (size_t)c->reg_count * sizeof(double), cudaMemcpyDeviceToDevice, c->stream);
        if (e == cudaSuccess) {
            k_rehash<<<nblk(c->reg_count, 256), 256, 0, c->stream>>>(nkeys, c->Wk, c->reg_count, ntable, (u32)(cap - 1)); LAUNCH(c);
            e = cudaGetLastError();
        }
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        cudaFree(ntable); cudaFree(nkeys); cudaFree(nll); cudaFree(nnp);
        return fail(c, BIC_ERR_CUDA, std::string("growing the family cache: ") + cudaGetErrorString(e));
    }
    long long keep = c->reg_count;
    cache_free(c);
    c->table = ntable; c->regkeys = nkeys; c->reg_ll = nll; c->reg_np = nnp;
    c->table_cap = cap; c->reg_cap = rcap; c->reg_count = keep;
    return BIC_OK;
}

int cache_clear(bic_ctx *c) {
    if (c->table) CU(cudaMemsetAsync(c->table, 0, c->table_cap * sizeof(u32), c->stream));
    c->reg_count = 0;
    c->lookups = 0;
    c->misses = 0;
    return BIC_OK;
}

double metric_penalty(bic_ctx *c, int metric) {
    if (metric == BIC_METRIC_BIC) return c->N_total > 0 ? 0.5 * log((double)c->N_total) : 0.0;
    if (metric == BIC_METRIC_AIC) return 1.0;
    return 0.0;
}

int header_reset(bic_ctx *c) {
    CU(cudaMemsetAsync(c->d_hdr, 0, sizeof(Header), c->stream));
    return BIC_OK;
}

int header_fetch(bic_ctx *c) {
    CU(cudaMemcpyAsync(c->h_hdr, c->d_hdr, sizeof(Header), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BIC_OK;
}

// Dynamic shared memory above 48 KB needs a per-device opt-in on the kernel; the largest size set
// so far is remembered per context (= per device) and per template instance.
template <int THREADS, bool GLOBAL, bool RANGE = false>
int launch_count(bic_ctx *c, const CountArgs &a, long long items, size_t smem) {
    size_t &attr_smem = c->attr_smem[RANGE ? (THREADS == 512 ? 7 : 8) : (THREADS == 256 ? 0 : THREADS == 512 ? 1 : 2) + (GLOBAL ? 3 : 0)];
    if (smem > 40 * 1024 && smem > attr_smem) {   // static + dynamic over 48 KB needs the opt-in
        CU(cudaFuncSetAttribute(k_count<THREADS, GLOBAL, RANGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    k_count<THREADS, GLOBAL, RANGE><<<(unsigned)items, THREADS, smem, c->stream>>>(a); LAUNCH(c);
    CU(cudaGetLastError());
    ++c->prof.count_launches;
    return BIC_OK;
}

// Row count over all ranks (ln N of the penalty) when the rows are sharded.
int refresh_ntotal(bic_ctx *c) {
    if (!c->ntotal_dirty) return BIC_OK;
    c->N_total = c->N;
    if (c->comm && c->comm_mode == BIC_SHARD_ROWS) {
        long long *d = nullptr;
        CU(cudaMalloc(&d, sizeof(long long)));
        CU(cudaMemcpyAsync(d, &c->N, sizeof(long long), cudaMemcpyHostToDevice, c->stream));
        int rc = g_nccl.AllReduce(d, d, 1, NCCL_INT64, NCCL_SUM, c->comm, c->stream);
        if (rc != 0) { cudaFree(d); return fail(c, BIC_ERR_NCCL, std::string("ncclAllReduce(N): ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?")); }
        CU(cudaMemcpyAsync(&c->N_total, d, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(d);
        // the summed count tables are uint32 on the wire and int32 at the API
        if (c->N_total >= (1ll << 31)) return fail(c, BIC_ERR_ARG, "row-sharded dataset holds 2^31 rows or more in total (int32 count tables)");
    }
    c->ntotal_dirty = false;
    return BIC_OK;
}

// Plan of one run_count(): row slices per family for each count-kernel class, and whether class 3
// is counted in shared-memory sub-range passes.  Pure host arithmetic (exported as bic_plan_slices
// so that the CPU test suite can exercise it).
//
// Slicing the rows has two purposes and a price.  (a) Few families -> enough CTAs to fill the GPU.
// (b) A dataset larger than L2 -> items run slice-major, so all resident CTAs sweep the same
// window of rows and each column segment comes from HBM once per launch instead of once per
// family.  The price: every slice sets up, zeroes, compacts and merges its shared-memory table
// into HBM with one L2 atomic per non-zero cell.  Versions a-h always cut into 32 MB windows; for
// 64 diabetes-shaped local-move candidates (5 GB of rows, a few dozen large-table families) that
// meant 162 slices whose merges cost 3x the counting itself.  Never under 64K rows per slice.
// CTA shape of class 0 (tables of <= 2048 cells).  256 threads x 24 KB on the uint8 path, x 48 KB when
// every column streams from the 2-bit packed copy.  Since the packed path builds its counter offsets
// in 32 bits (IDP.4A) lane replicas are no longer capped at cells * R <= 16383, and for batches whose
// class-0 tables are mostly mid-size (mean >= 256 cells: the alarm-shaped candidates, 250 - 2000
// cells) 512 threads sharing 96 KB (32 replicas up to 768 cells, 16 up to 1536; still 1024 threads
// per SM) cut the class-0 launch from 51.5 to 45.2 ms: by then the shared-memory data pipe was the
// limiter (91 %, half of its atomic wavefronts bank conflicts).  Tiny tables (pigs-shaped: <= 81
// cells) already run with 32 replicas and keep the small CTAs, and so do datasets of fewer than 2^20
// rows (sachs, 5000 rows: 0.085 -> 0.136 ms per class-0 launch with the wide shape).
// On the uint8 path (its 16-bit lanes keep the cells * R <= 16383 cap, and the in-flight rows need the L1) the
// same move is worth less: diabetes-shaped class-0 launch 1.135 ms at 256 x 24 KB, 1.094 / 1.080 / 1.109 ms at
// 512 x 48 / 64 / 96 KB, 1.57 ms at 1024 threads.  512 x 64 KB it is.
constexpr u32 CLASS0_WORDS_WIDE = 24576;
constexpr u32 CLASS0_WORDS_WIDE_U8 = 16384;
void class0_shape(bool all_packed, long long N, long long count0, long long cells0, const bic_ctx::Tuning &tune, int &threads, u32 &words) {
    threads = tune.class0_threads;
    words = all_packed ? tune.class0_words_packed : tune.class0_words;
    if (!tune.class0_explicit && tune.class0_wide && N >= (1ll << 20) && count0 > 0 && cells0 >= 256 * count0) {
        threads = 512;
        words = all_packed ? CLASS0_WORDS_WIDE : CLASS0_WORDS_WIDE_U8;
    }
}

void plan_count(const bic_plan_in_t &in, const bic_ctx::Tuning &tune, bic_plan_out_t &out) {
    const long long smax = std::max<long long>(1, in.N / 65536);
    // measured on B200 (profiles/): streaming loads, L2 atomics, rows per second one CTA counts, CTA set-up
    const double HBM_BPS = 6.0e12, RED_PER_S = 1.0e11, CTA_ROWS_PER_S = 5.0e9, CTA_SETUP_S = 4.0e-6;
    int c0t; u32 c0w;
    class0_shape(in.all_packed != 0, in.N, in.class_count[0], in.class_cells[0], tune, c0t, c0w);
    const long long resident[NCLASS] = {1024 / c0t, 2, 1, 4};   // CTAs of a class one SM holds (64 registers per thread; 192 KB tables)
    // class 3 in passes over shared-memory sub-ranges (k_count<1024, false, true>) when every table
    // of the launch fits range_passes sub-ranges and a slice holds at least 4 rows per cell
    const long long span = CLASS2_CELLS;
    const int P3gen = (int)((in.max_cells + span - 1) / span);
    const int P3 = in.passes3 > 0 ? (int)in.passes3 : P3gen;   // top split: as many as the generic cut; 16-bit counters: about half
    // One pass over a thread-block cluster whose CTAs share the table (k_count_cluster) when it fits
    // 8 x 192 KB of distributed shared memory; else sub-range passes; else (few rows) L2 atomics.
    int CL = 0;
    if (in.class_count[3] > 0 && tune.cluster && P3gen <= 8 && in.N >= 4ll * in.max_cells) {
        CL = 2;
        while (CL < P3gen) CL *= 2;
        if (tune.cluster_size > CL) CL = tune.cluster_size;
    }
    out.cluster = CL;
    const bool ranged = CL == 0 && in.class_count[3] > 0 && tune.range_passes > 0 && P3gen <= tune.range_passes &&
                        in.N >= 4ll * in.max_cells;
    out.ranged = ranged ? 1 : 0;
    out.passes = ranged ? P3 : 1;
    for (int k = 0; k < NCLASS; ++k) {
        const long long cnt = in.class_count[k];
        long long S = 1;
        if (cnt > 0) {
            // (a): the slice count that minimises rounds x (rows per CTA + set-up) + merge traffic.
            // Whole rounds matter when a class keeps one CTA per SM: 297 CTAs on 148 SMs take
            // three rounds, not two.
            const bool clu3 = k == 3 && CL > 0;
            const bool rng3 = k == 3 && (ranged || clu3);   // one CTA per SM, rows >> cells
            const long long ctas = (rng3 && !clu3 && in.items3 > 0) ? (long long)in.items3 : cnt * (clu3 ? CL : rng3 ? P3 : 1);
            const long long slots = (long long)in.sm_count * (rng3 ? 1 : resident[k]);
            const long long hi = std::min(rng3 ? std::min(smax, in.N / (4ll * in.max_cells)) : smax,
                                          std::max<long long>(1, 4 * slots / ctas));
            const double merge1 = (k == 3 && !rng3) ? 0.0 : (double)in.class_cells[k] / RED_PER_S;
            const double rowdiv = clu3 ? (double)CL : 1.0;   // the CTAs of a cluster share out the slice's rows
            double best = 0.0;
            for (long long s = 1; s <= hi; ++s) {
                const double rounds = (double)((ctas * s + slots - 1) / slots);
                const double t = rounds * (((double)in.N / (double)s / rowdiv) / CTA_ROWS_PER_S + CTA_SETUP_S) +
                                 ((s > 1 || in.tables_in_hbm) ? (double)s * merge1 : 0.0);
                if (s == 1 || t < best * 0.97) { best = t; S = s; }
            }
            if (!tune.slice_model) S = std::min(smax, ((long long)in.sm_count * 8 + cnt - 1) / cnt);
            // (b): L2 windows.  Families that run one after another (many more than the GPU holds
            // at once) need a window that stays in L2 until the last of them has passed: 32 MB.
            // Families that are all resident at once sweep the rows side by side anyway; wider
            // windows then mean fewer CTAs to set up, zero, compact and merge (pigs-shaped local
            // moves: 1.65 ms with 32 MB windows, 1.04-1.09 ms with 160-320 MB; diabetes-shaped
            // best at 96-160 MB).  In between (only the two ends are measured) the window shrinks with
            // the number of rounds a slice takes: 128 MB / rounds.
            // Taken only when the HBM traffic it saves (the class's algorithmic row bytes beyond
            // one pass over the dataset) outweighs the extra merges.
            // all-packed datasets stream a quarter of the bytes per row: the same L2 footprint is 4x the rows
            // (measured on the alarm-shaped step: 58.2 / 57.1 / 56.8 / 56.9 ms with 32 / 64 / 128 / 400 MB)
            const long long wmin = tune.l2_window * (in.all_packed ? 4 : 1);
            const long long win = !tune.slice_model ? tune.l2_window :
                ctas <= slots ? std::max(wmin, tune.l2_window_max) : std::max(wmin, tune.l2_window_max / 2 * slots / ctas);
            const long long s_l2 = ((long long)in.n * in.N + win - 1) / win;
            if (s_l2 > S && !rng3) {
                const double row_bytes = (double)in.class_alg_bytes[k] - 4.0 * (double)in.class_cells[k];
                const double saved = (row_bytes - (double)in.n * (double)in.N) / HBM_BPS;
                if (!tune.slice_model || saved > (double)(s_l2 - S) * merge1) S = s_l2;
            }
        }
        out.slices[k] = (int)std::max<long long>(1, std::min(smax, S));
    }
}

// Class 3 over a thread-block cluster: CL CTAs per (family, slice), one CTA per SM.
template <int THREADS>
int launch_count_cluster(bic_ctx *c, const CountArgs &a, long long items, int CL, size_t smem) {
    size_t &attr_smem = c->attr_smem[THREADS == 512 ? 9 : 10];
    if (smem > attr_smem) {
        CU(cudaFuncSetAttribute(k_count_cluster<THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_smem = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(items * CL));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CL;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CU(cudaLaunchKernelEx(&cfg, k_count_cluster<THREADS>, a, CL)); LAUNCH(c);
    ++c->prof.count_launches;
    return BIC_OK;
}

std::string nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"; }

// Exchange buffers of the fused reduce-scatter (row-sharded runs, world > 1).  Collective: every
// rank allocates world slots, exports the allocation with cudaIpcGetMemHandle, the handles travel
// with ncclAllGather and every peer's buffer is mapped with cudaIpcOpenMemHandle (which also enables
// peer access).  Any failure on any rank (all-reduce of the outcome) leaves every rank on the
// ncclAllReduce path.
int xchg_setup(bic_ctx *c) {
    if (c->xchg_state != 0) return BIC_OK;
    c->xchg_state = -1;
    if (!c->tune.push || (c->world < 2 && !c->tune.push_world1)) return BIC_OK;
    const int W = c->world;
    int ok = 1;
    c->xcap = (u64)(c->tune.xchg_mb << 20) / sizeof(u32);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&c->xchg, (size_t)W * c->xcap * sizeof(u32)) != cudaSuccess) { c->xchg = nullptr; ok = 0; }
    if (ok && cudaIpcGetMemHandle(&mine, c->xchg) != cudaSuccess) ok = 0;
    cudaGetLastError();
    // handles + outcome flags of every rank
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    uint8_t *d_all = nullptr;
    std::vector<uint8_t> h_all((size_t)W * rec, 0);
    CU(cudaMalloc(&d_all, (size_t)W * rec));
    memcpy(h_all.data() + (size_t)c->rank_id * rec, &mine, sizeof(mine));
    h_all[(size_t)c->rank_id * rec + sizeof(mine)] = (uint8_t)ok;
    CU(cudaMemcpyAsync(d_all + (size_t)c->rank_id * rec, h_all.data() + (size_t)c->rank_id * rec, rec, cudaMemcpyHostToDevice, c->stream));
    int rc = g_nccl.AllGather(d_all + (size_t)c->rank_id * rec, d_all, rec, NCCL_UINT8, c->comm, c->stream);
    if (rc != 0) { cudaFree(d_all); return fail(c, BIC_ERR_NCCL, "ncclAllGather(exchange-buffer handles): " + nccl_err(rc)); }
    CU(cudaMemcpyAsync(h_all.data(), d_all, (size_t)W * rec, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < W; ++r) ok = ok && h_all[(size_t)r * rec + sizeof(mine)];
    std::vector<u32 *> peers((size_t)W, nullptr);
    if (ok) {
        for (int r = 0; r < W && ok; ++r) {
            if (r == c->rank_id) { peers[r] = c->xchg; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, h_all.data() + (size_t)r * rec, sizeof(h));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            c->peer_open.push_back(ptr);
            peers[r] = (u32 *)ptr;
        }
    }
    // second round: did every rank map every peer?
    int *d_flag = nullptr;
    CU(cudaMalloc(&d_flag, sizeof(int)));
    CU(cudaMemcpyAsync(d_flag, &ok, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    rc = g_nccl.AllReduce(d_flag, d_flag, 1, NCCL_UINT32, NCCL_MIN, c->comm, c->stream);
    int all_ok = 0;
    if (rc == 0) {
        CU(cudaMemcpyAsync(&all_ok, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    cudaFree(d_flag);
    cudaFree(d_all);
    if (rc != 0) return fail(c, BIC_ERR_NCCL, "ncclAllReduce(exchange-buffer outcome): " + nccl_err(rc));
    if (!all_ok) {   // stay on the ncclAllReduce path
        for (void *ptr : c->peer_open) cudaIpcCloseMemHandle(ptr);
        c->peer_open.clear();
        if (c->xchg) { cudaFree(c->xchg); c->xchg = nullptr; }
        return BIC_OK;
    }
    CU(cudaMalloc(&c->d_peer, (size_t)W * sizeof(u32 *)));
    CU(cudaMemcpy(c->d_peer, peers.data(), (size_t)W * sizeof(u32 *), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&c->d_barrier, sizeof(int)));
    CU(cudaMemset(c->d_barrier, 0, sizeof(int)));
    c->xchg_state = 1;
    return BIC_OK;
}

void xchg_release(bic_ctx *c) {
    for (void *ptr : c->peer_open) cudaIpcCloseMemHandle(ptr);
    c->peer_open.clear();
    if (c->xchg) cudaFree(c->xchg);
    if (c->d_peer) cudaFree(c->d_peer);
    if (c->d_barrier) cudaFree(c->d_barrier);
    c->xchg = nullptr; c->d_peer = nullptr; c->d_barrier = nullptr;
    c->xchg_state = 0;
}

struct RunOut {          // where run_count() writes the family terms
    double *ll, *np;
    long long base;      // index of job 0
};

// Count (and reduce) `njobs` families described in c->cells_arr / c->class_jobs; the header in
// pinned memory holds the class counts.  keys/key_base select registry or key buffer.
// push: row-sharded with the fused reduce-scatter (c->owner / c->xoff describe the jobs).
int run_count(bic_ctx *c, const u64 *keys, long long key_base, long long njobs, long long max_jobs,
              RunOut out, bool want_tables, bool with_donors, bool push) {
    const Header &h = *c->h_hdr;
    const bool sharded = c->comm != nullptr && c->comm_mode == BIC_SHARD_ROWS;
    const u32 n_derived = with_donors ? h.n_derived : 0;
    const u32 max_cells = h.max_cells;
    u32 lvl_count[DERIVE_LEVELS];
    for (int l = 0; l < DERIVE_LEVELS; ++l) lvl_count[l] = with_donors ? h.lvl_count[l] : 0;
    const bool all_tables = want_tables || (sharded && !push) || n_derived > 0;

    // Row slices per family and how class 3 is counted: plan_count() above.
    NeedArgs na;
    na.all = all_tables ? 1 : 0;
    bic_plan_in_t pin;
    pin.sm_count = c->sm_count; pin.N = c->N; pin.n = c->n; pin.max_cells = h.max_cells; pin.tables_in_hbm = all_tables ? 1 : 0; pin.all_packed = c->all_packed ? 1 : 0;
    const bool c3_u16 = c->tune.c3_u16 && !c->tune.park_cells;
    // class 3 in passes: one work item per (family, pass it needs) when the decoded families are parked (k_range_items)
    const bool list3 = c->tune.park_meta && njobs <= (1ll << 22) && !c->tune.park_cells && c->tune.list3;
    pin.items3 = list3 ? (int)(c3_u16 ? h.sum_passes3_u16 : h.sum_passes3) : 0;
    pin.passes3 = c3_u16 ? (int)h.max_passes3_u16 : (c->tune.topsplit && !c->tune.park_cells) ? (int)h.max_passes3 : 0;
    for (int k = 0; k < NCLASS; ++k) {
        pin.class_count[k] = h.class_count[k];
        pin.class_cells[k] = (long long)h.class_cells[k];
        pin.class_alg_bytes[k] = (long long)h.alg_bytes[k];
    }
    bic_plan_out_t plan;
    plan_count(pin, c->tune, plan);
    const u32 span = CLASS2_CELLS;
    const int P3 = plan.passes;
    const bool ranged = plan.ranged != 0;
    const int CL = plan.cluster;
    bool any_table = all_tables;
    for (int k = 0; k < NCLASS; ++k) {
        na.S[k] = plan.slices[k];
        if (h.class_count[k] && (na.S[k] > 1 || k == 3)) any_table = true;
    }

    if (any_table) {
        CU(c->need.ensure((size_t)njobs * sizeof(u32)));
        CU(c->table_off.ensure((size_t)njobs * sizeof(u64)));
        k_table_need<<<nblk(njobs, 256), 256, 0, c->stream>>>(c->cells_arr.as<u32>(), (u32)njobs, na, c->need.as<u32>()); LAUNCH(c);
        TRY((scan_excl<u32, u64>(c, c->need.as<u32>(), njobs, c->table_off.as<u64>(), &c->d_hdr->table_cells, c->bsum64)));
        // no second host synchronisation: the first header already carries the sum of all cells,
        // an upper bound of what the scan assigns (exact when every table lives in HBM)
        size_t cells = (size_t)h.cells_all;
        CU(c->arena.ensure(std::max<size_t>(cells, 1) * sizeof(u32)));
        CU(cudaMemsetAsync(c->arena.p, 0, std::max<size_t>(cells, 1) * sizeof(u32), c->stream));
    }
    CU(c->done.ensure((size_t)njobs * sizeof(u32)));
    CU(cudaMemsetAsync(c->done.p, 0, (size_t)njobs * sizeof(u32), c->stream));

    CountArgs a = {};
    a.data = c->data; a.N = c->N; a.stride = c->stride; a.card = c->d_card; a.W64 = c->W64;
    a.data2 = c->data2; a.stride2 = c->stride2;
    a.keys = keys; a.key_base = key_base;
    a.arena = c->arena.as<u32>();
    a.need = any_table ? c->need.as<u32>() : nullptr;
    a.table_off = any_table ? c->table_off.as<u64>() : nullptr;
    a.done = c->done.as<u32>();
    a.ll_out = out.ll; a.np_out = out.np; a.out_base = out.base;
    a.reduce = sharded ? 0 : 1;
    a.donor = (with_donors && n_derived) ? c->donor.as<int>() : nullptr;
    a.bd_mode = c->cache_mode;
    a.p2_vec = c->tune.p2_vec;
    a.k30 = 1u << 30; a.k28 = 1u << 28; a.k26 = 1u << 26;
    a.tma = c->tune.tma ? 1 : 0;
    a.u8_narrow = c->tune.u8_narrow ? 1 : 0;
    a.swizzle = c->tune.swizzle ? 1 : 0;
    a.c3_u16 = c3_u16 ? 1 : 0;
    a.topsplit = (c->tune.topsplit && !c->tune.park_cells) ? 1 : 0;   // the parked cell indices follow the generic cut
    a.u8_two = c->tune.u8_two;
    a.p2_two = c->tune.p2_two;
    a.meta = nullptr;
    if (c->tune.park_meta && njobs <= (1ll << 22)) {   // decode every job's key once, not once per count CTA
        CU(c->meta.ensure((size_t)njobs * sizeof(FamMetaC)));
        k_decode_jobs<<<nblk(njobs, 128), 128, 0, c->stream>>>(keys, key_base, c->W64, c->d_card, njobs, c->meta.as<FamMetaC>()); LAUNCH(c);
        CU(cudaGetLastError());
        a.meta = c->meta.as<FamMetaC>();
    }
    a.iss = c->iss;
    a.push = push ? 1 : 0;
    a.rank = c->rank_id; a.world = c->world;
    a.owner = push ? c->owner.as<int>() : nullptr;
    a.xoff = push ? c->xoff.as<u64>() : nullptr;
    a.peer = push ? c->d_peer : nullptr;
    a.xcap = c->xcap;
    a.writeback = (push && n_derived > 0) ? 1 : 0;
    u32 class_count[NCLASS];
    u64 class_alg[NCLASS];
    for (int k = 0; k < NCLASS; ++k) { class_count[k] = h.class_count[k]; class_alg[k] = h.alg_bytes[k]; }

    for (int k = 0; k < NCLASS; ++k) {
        long long cnt = class_count[k];
        if (!cnt) continue;
        a.jobs = c->class_jobs.as<int>() + (long long)k * max_jobs;
        if (c->tune.sort_jobs && cnt > 1 && cnt <= (1ll << 20)) {   // most parents first (k_order_jobs)
            CU(c->jobs_sorted.ensure((size_t)max_jobs * NCLASS * sizeof(int)));
            int *sorted = c->jobs_sorted.as<int>() + (long long)k * max_jobs;
            k_order_jobs<<<1, 1024, 0, c->stream>>>(a.jobs, (int)cnt, keys, key_base, c->W64, sorted); LAUNCH(c);
            CU(cudaGetLastError());
            a.jobs = sorted;
        }
        a.S = na.S[k];
        a.njobs = (int)cnt;
        const bool clustered = k == 3 && CL > 0;
        a.P = (k == 3 && ranged) ? P3 : 1;   // launch-wide: the family with the most sub-ranges
        a.span = span;
        long long items = cnt * a.S * a.P;
        a.items3 = nullptr;
        a.nitems3 = 0;
        if (k == 3 && ranged && CL == 0 && list3 && a.meta && pin.items3 > 0) {
            CU(c->items3.ensure((size_t)pin.items3 * sizeof(int2)));
            k_range_items<<<1, 1024, 0, c->stream>>>(a.jobs, (int)cnt, a.meta, span, a.c3_u16, a.topsplit, c->items3.as<int2>(), (u32)pin.items3); LAUNCH(c);
            CU(cudaGetLastError());
            a.items3 = c->items3.as<int2>();
            a.nitems3 = pin.items3;
            items = (long long)pin.items3 * a.S;
        }
        if (items * (clustered ? CL : 1) > 0x7fffffffLL) return fail(c, BIC_ERR_ARG, "too many count work items in one launch");
        bic_ctx::EvPair ev = {nullptr, nullptr, k};
        if (c->prof_on) {   // CUDA events on the launching stream, one pair per count launch
            if (c->ev_pool.empty()) {
                CU(cudaEventCreate(&ev.a));
                CU(cudaEventCreate(&ev.b));
            } else {
                ev = c->ev_pool.back();
                ev.cls = k;
                c->ev_pool.pop_back();
            }
            CU(cudaEventRecord(ev.a, c->stream));
        }
        // shared memory per CTA: class 0 gets several times its largest table so that small tables
        // run with 32 or 16 bank-interleaved lane replicas (conflict-free atomics).
        int c0t; u32 c0w;
        class0_shape(c->all_packed, c->N, (long long)h.class_count[0], (long long)h.class_cells[0], c->tune, c0t, c0w);
        const u32 cap[NCLASS] = {c0w, CLASS1_CELLS, CLASS2_CELLS, 0};
        a.cap_words = cap[k];
        const u32 GLOBAL_STAGE = 8192;   // class 3 straight into HBM: shared memory only stages the final reduce
        const u32 clwords = clustered ? (u32)((max_cells + CL - 1) / CL) : 0;
        a.cellbuf = nullptr;
        if (k == 3 && ranged && c->tune.park_cells && cnt <= 65535) {   // compute every row's cell once, the passes only compare and increment
            const size_t bytes = (size_t)cnt * (size_t)c->stride * sizeof(u32);
            if (bytes <= ((size_t)c->tune.cells_max_mb << 20)) {
                CU(c->cellbuf.ensure(bytes));
                a.cellbuf = c->cellbuf.as<u32>();
                const long long nvec = (c->N + 15) >> 4;
                dim3 grid((unsigned)std::min<long long>((nvec + 255) / 256, std::max<long long>(1, (long long)c->sm_count * 8 / cnt)), (unsigned)cnt);
                k_cells<<<grid, 256, 0, c->stream>>>(a); LAUNCH(c);
                CU(cudaGetLastError());
            }
        }
        a.stage_words = k == 3 ? (clustered ? clwords : ranged ? span : GLOBAL_STAGE) : cap[k];
        const size_t ring256 = a.tma ? tma_ring_bytes(256) : 0, ring512 = a.tma ? tma_ring_bytes(512) : 0;   // TMA staging ring behind the table
        // Tiers (all-packed datasets in the wide class-0 shape): the tables above the replica reach of a 96 KB CTA
        // are counted by a second launch over the same class list with 1024 threads x 192 KB (32 replicas up to
        // 1536 cells, 16 up to 3072); a CTA whose family belongs to the other launch returns at once.
        const bool tiers = c->all_packed && c0t == 512 && c0w == CLASS0_WORDS_WIDE && !a.tma && (c->tune.tier0 || c->tune.tier1);
        a.tier_lo = a.tier_hi = 0;
        if (k == 0 && tiers && c->tune.tier0) {
            a.tier_lo = 0; a.tier_hi = c->tune.tier0;
            TRY((launch_count<512, false>(c, a, items, cap[0] * sizeof(u32))));
            a.tier_lo = c->tune.tier0 + 1; a.tier_hi = 0xffffffffu;
            a.cap_words = a.stage_words = CLASS2_CELLS;
            TRY((launch_count<1024, false>(c, a, items, CLASS2_CELLS * sizeof(u32))));
            a.tier_lo = a.tier_hi = 0;
        } else {
        if (k == 0 && c0t == 256) TRY((launch_count<256, false>(c, a, items, cap[0] * sizeof(u32) + ring256)));
        if (k == 0 && c0t == 512) TRY((launch_count<512, false>(c, a, items, cap[0] * sizeof(u32) + ring512)));
        if (k == 0 && c0t == 1024) TRY((launch_count<1024, false>(c, a, items, cap[0] * sizeof(u32))));
        }
        const int c1t = c->tune.class1_threads;
        if (k == 1 && tiers && c->tune.tier1) {
            a.tier_lo = 0; a.tier_hi = c->tune.tier1;
            a.cap_words = a.stage_words = CLASS2_CELLS;
            TRY((launch_count<1024, false>(c, a, items, CLASS2_CELLS * sizeof(u32))));
            a.tier_lo = c->tune.tier1 + 1; a.tier_hi = 0xffffffffu;
            a.cap_words = a.stage_words = cap[1];
            TRY((launch_count<512, false>(c, a, items, cap[1] * sizeof(u32))));
            a.tier_lo = a.tier_hi = 0;
        } else {
        if (k == 1 && c1t == 256) TRY((launch_count<256, false>(c, a, items, cap[1] * sizeof(u32) + ring256)));
        if (k == 1 && c1t == 512) TRY((launch_count<512, false>(c, a, items, cap[1] * sizeof(u32) + ring512)));
        if (k == 1 && c1t == 1024) TRY((launch_count<1024, false>(c, a, items, cap[1] * sizeof(u32))));
        }
        const bool wide = c->tune.class2_threads == 1024;
        if (k == 2 && !wide) TRY((launch_count<512, false>(c, a, items, cap[2] * sizeof(u32))));
        if (k == 2 && wide) TRY((launch_count<1024, false>(c, a, items, cap[2] * sizeof(u32))));
        if (k == 3 && clustered && c->tune.cluster_threads == 512) TRY((launch_count_cluster<512>(c, a, items, CL, clwords * sizeof(u32))));
        if (k == 3 && clustered && c->tune.cluster_threads != 512) TRY((launch_count_cluster<1024>(c, a, items, CL, clwords * sizeof(u32))));
        if (k == 3 && !clustered && !ranged) TRY((launch_count<256, true>(c, a, items, GLOBAL_STAGE * sizeof(u32))));
        if (k == 3 && ranged && !wide) TRY((launch_count<512, false, true>(c, a, items, span * sizeof(u32))));
        if (k == 3 && ranged && wide) TRY((launch_count<1024, false, true>(c, a, items, span * sizeof(u32))));
        if (c->prof_on) {
            CU(cudaEventRecord(ev.b, c->stream));
            c->ev_used.push_back(ev);
        }
        c->prof.class_launches[k] += 1;
        c->prof.class_families[k] += cnt;
        c->prof.class_alg_bytes[k] += (long long)class_alg[k];
        c->prof.alg_bytes += (long long)class_alg[k];
    }
    long long counted = 0;   // jobs of this rank (all new families unless they are sharded over the ranks)
    for (int k = 0; k < NCLASS; ++k) counted += class_count[k];
    c->prof.families_counted += counted;
    c->prof.rows_counted += counted * c->N;
    c->prof.families_derived += n_derived;

    if (sharded) {
        bic_ctx::EvPair ev = {nullptr, nullptr, -1};   // cls -1: the exchange step (collective + owner reduce)
        if (c->prof_on) {
            if (c->ev_pool.empty()) { CU(cudaEventCreate(&ev.a)); CU(cudaEventCreate(&ev.b)); }
            else { ev = c->ev_pool.back(); ev.cls = -1; c->ev_pool.pop_back(); }
            CU(cudaEventRecord(ev.a, c->stream));
        }
        if (push) {
            // The count kernels have already stored every partial table into its owner's exchange
            // buffer.  A 4-byte all-reduce orders "all ranks' count kernels are complete" before the
            // owners read their slots; the owner then sums the world slots inside the fp64 reduce.
            int rc = g_nccl.AllReduce(c->d_barrier, c->d_barrier, 1, NCCL_UINT32, NCCL_SUM, c->comm, c->stream);
            if (rc != 0) return fail(c, BIC_ERR_NCCL, "ncclAllReduce(push barrier): " + nccl_err(rc));
            long long counted_cells = 0;
            for (int k = 0; k < NCLASS; ++k) counted_cells += (long long)h.class_cells[k];
            c->prof.exchange_bytes += counted_cells * 4 * (c->world - 1) / c->world;   // what this rank stores into peers' buffers
            ++c->prof.exchange_fused;
            k_sum_slots<<<c->sm_count * 4, 256, 0, c->stream>>>(c->xchg, c->xcap, c->world, c->rank_id, c->d_hdr); LAUNCH(c);
        } else {
            size_t cells = (size_t)h.cells_all;   // all tables live in HBM when sharded: exact
            if (cells) {
                int rc = g_nccl.AllReduce(c->arena.p, c->arena.p, cells, NCCL_UINT32, NCCL_SUM, c->comm, c->stream);
                if (rc != 0) return fail(c, BIC_ERR_NCCL, "ncclAllReduce(count tables): " + nccl_err(rc));
            }
            c->prof.exchange_bytes += (long long)cells * 4;
            ++c->prof.exchange_nccl;
        }
        k_reduce_tables<256><<<(unsigned)njobs, 256, 0, c->stream>>>(a, (int)njobs); LAUNCH(c);
        CU(cudaGetLastError());
        if (c->prof_on) {
            CU(cudaEventRecord(ev.b, c->stream));
            c->ev_used.push_back(ev);
        }
    }
    if (n_derived) {   // tables of the counted families are complete: marginalise, most parents first
        // chunks per family: one per 4096 donor cells, as many as the largest table of the batch can need
        const int dch = (int)std::min<u32>(DERIVE_CHUNKS, std::max<u32>(1u, max_cells / 4096u));
        long long first = 0;
        for (int l = DERIVE_LEVELS - 1; l >= 0; --l) {
            if (!lvl_count[l]) continue;
            k_derive<256><<<lvl_count[l] * dch, 256, 0, c->stream>>>(a, c->derived_sorted.as<int>() + first, dch,
                                                                     c->cells_arr.as<u32>()); LAUNCH(c);
            first += lvl_count[l];
        }
        CU(cudaGetLastError());
    }
    return BIC_OK;
}

int check_header_err(bic_ctx *c) {
    u32 e = c->h_hdr->err;
    if (e & 2u) return fail(c, BIC_ERR_BAD_FAMILY, "a parent index is out of range or equals its node");
    if (e & 1u) return fail(c, BIC_ERR_TABLE_TOO_LARGE, "a family's count table q*r exceeds 2^28 cells");
    if (e & 4u) return fail(c, BIC_ERR_ARG, "counts_off does not match q*r of a family");
    return BIC_OK;
}

// keys of T instances sit in c->keybuf: look them up, insert + count the unseen families.
int resolve_instances(bic_ctx *c, long long T, int n_per_dag, bool no_derive) {
    const bool famshard = c->comm != nullptr && c->comm_mode == BIC_SHARD_FAMILIES && c->world > 1;
    const bool rowshard = c->comm != nullptr && c->comm_mode == BIC_SHARD_ROWS;
    // Every allocation happens before k_probe: from there to the end a failure would leave PENDING
    // entries or ids without terms in the table, so any non-OK return below drops the cache.
    TRY(cache_ensure(c, T));
    CU(c->inst.ensure((size_t)T * sizeof(int)));
    CU(c->flag.ensure((size_t)T * sizeof(u32)));
    CU(c->rank.ensure((size_t)T * sizeof(u32)));
    CU(c->cells_arr.ensure((size_t)T * sizeof(u32)));
    CU(c->class_jobs.ensure((size_t)T * NCLASS * sizeof(int)));
    CU(c->donor.ensure((size_t)T * sizeof(int)));
    CU(c->donor_best.ensure((size_t)T * sizeof(u64)));   // (joint cells, index) of the cheapest donor announced so far
    CU(c->derived_list.ensure((size_t)T * sizeof(int)));
    CU(c->derived_sorted.ensure((size_t)T * sizeof(int)));
    CU(c->bsum32.ensure((size_t)((T + SCAN_CHUNK - 1) / SCAN_CHUNK + 1) * sizeof(u32)));
    const bool multi = c->world > 1 || c->tune.push_world1;
    if (rowshard && multi) {
        TRY(xchg_setup(c));   // collective, first call only
        CU(c->owner.ensure((size_t)T * sizeof(int)));
        CU(c->xoff.ensure((size_t)T * sizeof(u64)));
    }
    const bool can_push = rowshard && multi && c->xchg_state == 1;
    struct Guard {
        bic_ctx *c;
        bool armed;
        ~Guard() { if (armed) cache_clear(c); }
    } guard{c, true};

    unsigned g = nblk(T, 256);
    k_probe<<<g, 256, 0, c->stream>>>(c->keybuf.as<u64>(), c->Wk, T, n_per_dag, c->dag_bad.as<uint8_t>(), c->table,
                                      (u32)(c->table_cap - 1), c->regkeys, c->inst.as<int>()); LAUNCH(c);
    k_owner_flags<<<g, 256, 0, c->stream>>>(c->inst.as<int>(), c->table, T, c->flag.as<u32>()); LAUNCH(c);
    TRY((scan_excl<u32, u32>(c, c->flag.as<u32>(), T, c->rank.as<u32>(), &c->d_hdr->f_new, c->bsum32)));
    k_finalize<<<g, 256, 0, c->stream>>>(c->keybuf.as<u64>(), c->Wk, T, c->inst.as<int>(), c->flag.as<u32>(),
                                         c->rank.as<u32>(), c->reg_count, c->regkeys, c->table); LAUNCH(c);
    // New families: find superset donors (large datasets only), then describe / classify.  The
    // decision to derive is taken from values every rank agrees on (row-sharded ranks hold shards
    // that differ by a row: N_total / world, not the local N).
    const long long rows_ref = rowshard ? c->N_total / std::max(1, c->world) : c->N;
    const int derive = (c->tune.derive && !no_derive && rows_ref >= c->tune.derive_min_rows) ? 1 : 0;
    if (derive) {
        const int aw = famshard ? c->world : 1;
        long long fmax = T;   // upper bound of f_new
        if (famshard) {       // one more header fetch buys an all-reduce over f_new instead of T entries
            TRY(header_fetch(c));
            fmax = c->h_hdr->f_new;
        }
        if (fmax > 0) {
            CU(cudaMemsetAsync(c->donor_best.p, 0xff, (size_t)fmax * sizeof(u64), c->stream));
            k_announce<<<nblk(((fmax + aw - 1) / aw) * 32, 256), 256, 0, c->stream>>>(c->regkeys, c->W64, c->reg_count, c->d_hdr, c->table,
                                                                              (u32)(c->table_cap - 1), c->d_card, c->donor_best.as<u64>(),
                                                                              famshard ? c->rank_id : 0, aw); LAUNCH(c);
            if (famshard) {   // every rank announced for 1/world of the donors: combine the minima
                int e = g_nccl.AllReduce(c->donor_best.p, c->donor_best.p, (size_t)fmax, NCCL_UINT64, NCCL_MIN, c->comm, c->stream);
                if (e != 0) return fail(c, BIC_ERR_NCCL, "ncclAllReduce(donor search): " + nccl_err(e));
            }
        }
        k_pick_donor<<<g, 256, 0, c->stream>>>(c->donor_best.as<u64>(), c->d_hdr, derive, c->donor.as<int>()); LAUNCH(c);
    }
    k_describe_new<<<g, 256, 0, c->stream>>>(c->regkeys, c->W64, c->reg_count, c->d_card, c->N, (u32)T, c->d_hdr,
                                             derive ? c->donor.as<int>() : nullptr, c->cells_arr.as<u32>(), c->class_jobs.as<int>(),
                                             c->derived_list.as<int>(), c->rank_id, (famshard || can_push) ? c->world : 1,
                                             famshard ? 1 : 0, can_push ? c->owner.as<int>() : nullptr); LAUNCH(c);
    if (derive) {
        k_group_derived<<<g, 256, 0, c->stream>>>(c->regkeys, c->W64, c->reg_count, c->d_hdr, c->derived_list.as<int>(),
                                                  c->derived_sorted.as<int>()); LAUNCH(c);
    }
    if (can_push) {
        k_owner_offsets<<<c->world, 1024, 0, c->stream>>>(c->cells_arr.as<u32>(), c->owner.as<int>(),
                                                          derive ? c->donor.as<int>() : nullptr, c->d_hdr, c->xoff.as<u64>()); LAUNCH(c);
    }
    CU(cudaGetLastError());
    TRY(header_fetch(c));
    long long f_new = c->h_hdr->f_new;
    c->lookups += T;
    c->misses += f_new;
    TRY(check_header_err(c));   // new ids were published without scores: the guard drops the whole cache
    if (f_new) {
        // push only when every rank's owned tables fit a slot of the exchange buffer (same header on every rank)
        const bool push = can_push && c->h_hdr->owned_max <= c->xcap;
        const bool staged = famshard || push;   // terms of the families other ranks own arrive as x + 0 + ... + 0
        RunOut out = {c->reg_ll, c->reg_np, c->reg_count};
        if (staged) {
            CU(c->terms.ensure((size_t)f_new * 2 * sizeof(double)));
            CU(cudaMemsetAsync(c->terms.p, 0, (size_t)f_new * 2 * sizeof(double), c->stream));
            out = RunOut{c->terms.as<double>(), c->terms.as<double>() + f_new, 0};
        }
        TRY(run_count(c, c->regkeys, c->reg_count, f_new, T, out, false, true, push));
        if (staged) {
            int e = g_nccl.AllReduce(c->terms.p, c->terms.p, (size_t)f_new * 2, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
            if (e != 0) return fail(c, BIC_ERR_NCCL, "ncclAllReduce(family terms): " + nccl_err(e));
            CU(cudaMemcpyAsync(c->reg_ll + c->reg_count, c->terms.p, (size_t)f_new * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(c->reg_np + c->reg_count, c->terms.as<double>() + f_new, (size_t)f_new * sizeof(double),
                               cudaMemcpyDeviceToDevice, c->stream));
        }
        c->reg_count += f_new;
    }
    guard.armed = false;
    return BIC_OK;
}

// Resolve the CUDA-event pairs recorded around the count launches of this call.
int finish_call(bic_ctx *c) {
    CU(cudaStreamSynchronize(c->stream));
    for (auto &p : c->ev_used) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            if (p.cls < 0) {
                c->prof.exchange_ms += ms;
            } else {
                c->prof.count_ms += ms;
                c->prof.class_ms[p.cls] += ms;
            }
        }
        c->ev_pool.push_back(p);
    }
    c->ev_used.clear();
    return BIC_OK;
}

// Host or device input -> device pointer (staged when host).
template <typename T>
int stage_in(bic_ctx *c, const T *src, size_t count, int flags, DevBuf &buf, const T **out) {
    if (flags & BIC_FLAG_DEVICE_PTRS) { *out = src; return BIC_OK; }
    CU(buf.ensure(std::max<size_t>(count, 1) * sizeof(T)));
    if (count) CU(cudaMemcpyAsync(buf.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *out = buf.as<T>();
    return BIC_OK;
}

int begin_call(bic_ctx *c, int metric, bool need_metric) {
    if (!c->data) return fail(c, BIC_ERR_NO_DATASET, "no dataset: call bic_set_dataset first");
    if (need_metric && (metric < BIC_METRIC_BIC || metric > BIC_METRIC_K2)) return fail(c, BIC_ERR_ARG, "unknown metric");
    CU(cudaSetDevice(c->device));
    TRY(refresh_ntotal(c));
    if (need_metric) {   // the cache holds one kind of family term at a time
        int mode = metric == BIC_METRIC_BDE ? 1 : metric == BIC_METRIC_K2 ? 2 : 0;
        if (mode != c->cache_mode) {
            TRY(cache_clear(c));
            c->cache_mode = mode;
        }
    }
    return BIC_OK;
}

long long sub_batch_dags(bic_ctx *c) {
    const long long max_inst = 1ll << 22;
    return std::max<long long>(1, max_inst / c->n);
}

enum DagFormat { FMT_ADJ, FMT_CSR, FMT_WIRE, FMT_WIRE16 };

// ewords: FMT_WIRE16 only, 32-bit edge words per vertex.
int score_dags(bic_ctx *c, DagFormat fmt, const void *p0, const void *p1, int64_t B, int metric, double *out,
               int64_t *n_invalid, int flags, int ewords = 1) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    TRY(begin_call(c, metric, true));
    if (B < 0 || (B > 0 && (!p0 || !out))) return fail(c, BIC_ERR_ARG, "null pointer or negative batch size");
    if (fmt == FMT_WIRE && c->n > 32) return fail(c, BIC_ERR_ARG, "bic_score_dags_wire supports n <= 32; use bic_score_dags_wire16");
    if (fmt == FMT_WIRE16 && ewords < (c->n + 31) / 32) return fail(c, BIC_ERR_ARG, "ewords < ceil(n / 32)");
    const bool famshard = c->comm != nullptr && c->comm_mode == BIC_SHARD_FAMILIES && c->world > 1;
    // BIC_FLAG_LOCAL_BATCH (family sharding): the arrays hold THIS rank's B DAGs (same B on every
    // rank).  Keys are built and cycles checked locally, the keys travel with ncclAllGather over
    // NVLink (16 bytes per family instead of n bytes of adjacency through every rank's PCIe link),
    // the union is deduplicated identically on every rank and only the local scores come back.
    const bool local = (flags & BIC_FLAG_LOCAL_BATCH) != 0;
    if (local && !famshard) return fail(c, BIC_ERR_ARG, "BIC_FLAG_LOCAL_BATCH needs family sharding over more than one rank");
    if (local && fmt == FMT_CSR && !(flags & BIC_FLAG_DEVICE_PTRS)) return fail(c, BIC_ERR_ARG, "BIC_FLAG_LOCAL_BATCH: CSR input must be device-resident");
    if (flags & BIC_FLAG_NO_CACHE) TRY(cache_clear(c));
    const int n = c->n;
    const int W = local ? c->world : 1, R = local ? c->rank_id : 0;
    const double pen = metric_penalty(c, metric);
    const bool dev = (flags & BIC_FLAG_DEVICE_PTRS) != 0;
    long long invalid = 0;
    // Small warm batch (the reference's one-DAG-per-call usage): one launch, one synchronisation.
    if (fmt == FMT_ADJ && c->fast_ok && c->tune.fast_small && !c->comm && c->reg_count > 0 && c->W64 == 1 && B > 0 &&
        B * n <= 8192) {
        TRY(header_reset(c));
        const uint8_t *adj = nullptr;
        TRY(stage_in(c, (const uint8_t *)p0, (size_t)B * n * n, flags, c->in_stage, &adj));
        double *d_out = out;
        if (!dev) {
            CU(c->out_stage.ensure((size_t)B * sizeof(double)));
            d_out = c->out_stage.as<double>();
        }
        k_score_small<<<nblk(B, SMALL_WARPS), SMALL_WARPS * 32, 0, c->stream>>>(
            adj, B, n, c->table, (u32)(c->table_cap - 1), c->regkeys, c->reg_ll, c->reg_np, pen,
            (flags & BIC_FLAG_NO_CYCLE_CHECK) ? 0 : 1, d_out, c->d_hdr); LAUNCH(c);
        CU(cudaGetLastError());
        if (!dev) CU(cudaMemcpyAsync(out, d_out, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TRY(header_fetch(c));
        if (!(c->h_hdr->err & 8u)) {
            c->lookups += B * n;
            if (n_invalid) *n_invalid = c->h_hdr->n_invalid;
            return finish_call(c);
        }
        c->fast_ok = false;   // a family is not cached yet: the general pipeline inserts and counts it
    }
    const long long misses0 = c->misses;
    const long long Bs = std::max<long long>(1, sub_batch_dags(c) / W);
    std::vector<long long> h_off;   // CSR offsets are needed on the host to slice a host batch
    // B == 0 still takes part in the collectives of a local-batch call
    for (long long b0 = 0; b0 < B || (local && b0 == 0); b0 += Bs) {
        const long long Bc = std::min<long long>(Bs, B - b0);   // this rank's DAGs of the sub-batch
        const long long Tl = Bc * n, T = Tl * W;                  // local / global family instances
        CU(c->keybuf.ensure((size_t)std::max<long long>(T, 1) * c->Wk * sizeof(u64)));
        CU(c->dag_bad.ensure((size_t)std::max<long long>(Bc * W, 1)));
        u64 *lkeys = c->keybuf.as<u64>() + (size_t)R * Tl * c->Wk;   // this rank's part of the (global) key buffer
        uint8_t *lbad = c->dag_bad.as<uint8_t>() + (size_t)R * Bc;
        CU(cudaMemsetAsync(c->dag_bad.p, 0, (size_t)std::max<long long>(Bc * W, 1), c->stream));
        TRY(header_reset(c));
        if (Bc > 0) {
        if (fmt == FMT_ADJ) {
            const uint8_t *adj = nullptr;
            TRY(stage_in(c, (const uint8_t *)p0 + b0 * (long long)n * n, (size_t)Bc * n * n, flags, c->in_stage, &adj));
            k_keys_adj<<<nblk(Tl, 256), 256, 0, c->stream>>>(adj, Bc, n, c->W64, lkeys, lbad); LAUNCH(c);
        } else if (fmt == FMT_CSR) {
            const long long *off = (const long long *)p0 + b0 * n;
            const int *par = (const int *)p1;
            const long long *d_off = nullptr;
            const int *d_par = nullptr;
            if (dev) {
                d_off = off;     // absolute offsets into the caller's device array
                d_par = par;
            } else {
                long long e0 = off[0], e1 = off[Tl];
                if (e1 < e0) return fail(c, BIC_ERR_ARG, "CSR offsets are not monotone");
                h_off.resize((size_t)Tl + 1);
                for (long long i = 0; i <= Tl; ++i) h_off[(size_t)i] = off[i] - e0;
                TRY(stage_in(c, h_off.data(), (size_t)Tl + 1, 0, c->in_stage, &d_off));
                TRY(stage_in(c, par + e0, (size_t)(e1 - e0), 0, c->in_stage2, &d_par));
                CU(cudaStreamSynchronize(c->stream));   // h_off is reused by the next sub-batch
            }
            k_keys_csr<<<nblk(Tl, 256), 256, 0, c->stream>>>(d_off, d_par, nullptr, Tl, n, c->W64, lkeys, lbad, c->d_hdr); LAUNCH(c);
        } else if (fmt == FMT_WIRE) {
            const uint8_t *lab = nullptr;
            const u32 *eb = nullptr;
            TRY(stage_in(c, (const uint8_t *)p0 + b0 * n, (size_t)Bc * n, flags, c->in_stage, &lab));
            TRY(stage_in(c, (const u32 *)p1 + b0 * n, (size_t)Bc * n, flags, c->in_stage2, &eb));
            k_keys_wire<uint8_t><<<nblk(Bc, 128), 128, 0, c->stream>>>(lab, eb, Bc, n, lkeys, lbad); LAUNCH(c);
        } else {
            const uint16_t *lab = nullptr;
            const u32 *eb = nullptr;
            TRY(stage_in(c, (const uint16_t *)p0 + b0 * n, (size_t)Bc * n, flags, c->in_stage, &lab));
            TRY(stage_in(c, (const u32 *)p1 + b0 * n * (long long)ewords, (size_t)Bc * n * ewords, flags, c->in_stage2, &eb));
            if (n <= 32 && ewords == 1) { k_keys_wire<uint16_t><<<nblk(Bc, 128), 128, 0, c->stream>>>(lab, eb, Bc, n, lkeys, lbad); LAUNCH(c); }
            else { k_keys_wire_wide<<<(unsigned)Bc, WIRE_WIDE_THREADS, 0, c->stream>>>(lab, eb, Bc, n, ewords, c->W64, lkeys, lbad); LAUNCH(c); }
        }
        if ((flags & BIC_FLAG_NO_CYCLE_CHECK) || fmt == FMT_WIRE || fmt == FMT_WIRE16) {
            k_count_bad<<<nblk(Bc, 256), 256, 0, c->stream>>>(lbad, Bc, c->d_hdr); LAUNCH(c);
        } else {
            if (n > 128) {
                const size_t mb = (size_t)n * c->W64 * sizeof(u64);
                const int staged = mb <= ACYC_SMEM_MAX ? 1 : 0;
                k_acyclic_wide<<<(unsigned)Bc, ACYC_WIDE_THREADS, staged ? mb : 0, c->stream>>>(lkeys, Bc, n, c->W64, lbad, c->d_hdr, staged);
            } else {
                k_acyclic_warp<<<nblk(Bc, ACYC_WARPS), ACYC_WARPS * 32, 0, c->stream>>>(lkeys, Bc, n, c->W64, lbad, c->d_hdr);
            }
            LAUNCH(c);
        }
        CU(cudaGetLastError());
        }
        if (local) {   // in-place all-gather: every rank's keys / reject flags at its own offset
            if (Tl > 0) {
                int e1 = g_nccl.AllGather(lkeys, c->keybuf.p, (size_t)Tl * c->Wk * sizeof(u64), NCCL_UINT8, c->comm, c->stream);
                int e2 = g_nccl.AllGather(lbad, c->dag_bad.p, (size_t)Bc, NCCL_UINT8, c->comm, c->stream);
                if (e1 != 0 || e2 != 0) return fail(c, BIC_ERR_NCCL, "ncclAllGather(candidate keys): " + nccl_err(e1 ? e1 : e2));
            }
        }
        if (T > 0) TRY(resolve_instances(c, T, n, (flags & BIC_FLAG_NO_DERIVE) != 0));
        if (Bc <= 0) break;
        invalid += c->h_hdr->n_invalid;
        double *d_out = out + b0;
        if (!dev) {
            CU(c->out_stage.ensure((size_t)Bc * sizeof(double)));
            d_out = c->out_stage.as<double>();
        }
        const int *linst = c->inst.as<int>() + (size_t)R * Tl;
        if (n >= 64)
            k_gather_dags_warp<<<nblk(Bc * 32, 128), 128, 0, c->stream>>>(linst, c->table, Bc, n, lbad, c->reg_ll, c->reg_np, pen, d_out);
        else
            k_gather_dags<<<nblk(Bc, 128), 128, 0, c->stream>>>(linst, c->table, Bc, n, lbad, c->reg_ll, c->reg_np, pen, d_out);
        LAUNCH(c);
        CU(cudaGetLastError());
        if (!dev) CU(cudaMemcpyAsync(out + b0, d_out, (size_t)Bc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TRY(finish_call(c));
    }
    if (n_invalid) *n_invalid = invalid;
    c->fast_ok = (c->misses == misses0);   // everything was cached: the next small batch may take the short cut
    return BIC_OK;
}

}  // namespace

// ================================================================================ C ABI
extern "C" {

int bic_range_plan(uint32_t cells, int32_t k, uint32_t rad0, int32_t counters16, uint32_t *span, uint32_t *passes,
                   uint32_t *states_per_pass) {
    if (!span || !passes || !states_per_pass || cells == 0 || k < 0 || rad0 == 0) return BIC_ERR_ARG;
    const RangePlan p = (counters16 && k <= 6) ? range_plan(cells, k, rad0, 2u * CLASS2_CELLS, false)
                                               : range_plan(cells, k, rad0, CLASS2_CELLS, true);
    *span = p.span;
    *passes = p.passes;
    *states_per_pass = p.ns;
    return BIC_OK;
}

int bic_plan_slices(const bic_plan_in_t *in, bic_plan_out_t *out) {
    if (!in || !out || in->sm_count <= 0 || in->N <= 0 || in->n <= 0 || in->max_cells < 0) return BIC_ERR_ARG;
    for (int k = 0; k < NCLASS; ++k)
        if (in->class_count[k] < 0 || in->class_cells[k] < 0 || in->class_alg_bytes[k] < 0) return BIC_ERR_ARG;
    bic_ctx::Tuning tune;
    tune.from_env();
    plan_count(*in, tune, *out);
    return BIC_OK;
}

int bic_version(void) { return BICGPU_VERSION; }

#define BIC_STR2(x) #x
#define BIC_STR(x) BIC_STR2(x)
const char *bic_build_info(void) {
    return "libbicgpu ABI " BIC_STR(BICGPU_VERSION) ", sm_100a only, nvcc " BIC_STR(__CUDACC_VER_MAJOR__) "." BIC_STR(__CUDACC_VER_MINOR__) "."
           BIC_STR(__CUDACC_VER_BUILD__) ", built " __DATE__ " " __TIME__;
}

const char *bic_last_error(const bic_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int bic_create(bic_ctx **out, int device) {
    bic_ctx *c = nullptr;
    if (!out) return fail(c, BIC_ERR_ARG, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(c, BIC_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libbicgpu has no CPU fallback)");
    if (device < 0 || device >= count) return fail(c, BIC_ERR_ARG, "device index out of range");
    cudaDeviceProp prop;
    CU(cudaSetDevice(device));
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(c, BIC_ERR_CUDA, std::string("libbicgpu is built for sm_100a only; device is ") + prop.name);
    bic_ctx *ctx = new bic_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    c = ctx;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaMalloc(&ctx->d_hdr, sizeof(Header)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_hdr, sizeof(Header)) != cudaSuccess) {
        g_create_error = std::string("context allocation failed: ") + cudaGetErrorString(cudaGetLastError());
        delete ctx;
        return BIC_ERR_CUDA;
    }
    ctx->stream = ctx->own_stream;
    ctx->tune.from_env();
    *out = ctx;
    return BIC_OK;
}

int bic_destroy(bic_ctx *c) {
    if (!c) return BIC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    xchg_release(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->ev_wait) cudaEventDestroy(c->ev_wait);
    cache_free(c);
    DevBuf *bufs[] = {&c->keybuf, &c->inst, &c->flag, &c->rank, &c->bsum32, &c->bsum64, &c->cells_arr, &c->class_jobs,
                      &c->need, &c->table_off, &c->done, &c->arena, &c->dag_bad, &c->in_stage, &c->in_stage2,
                      &c->in_stage3, &c->in_stage4, &c->out_stage, &c->tmp_ll, &c->donor, &c->donor_best,
                      &c->derived_list, &c->derived_sorted, &c->owner, &c->xoff, &c->fp_buf, &c->terms, &c->gkeys, &c->gbad, &c->cellbuf, &c->meta, &c->jobs_sorted, &c->items3};
    for (DevBuf *b : bufs) b->release();
    if (c->data) cudaFree(c->data);
    if (c->data2) cudaFree(c->data2);
    if (c->d_card) cudaFree(c->d_card);
    if (c->d_hdr) cudaFree(c->d_hdr);
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    for (auto &p : c->ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return BIC_OK;
}

int bic_set_stream(bic_ctx *c, void *cuda_stream) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return BIC_OK;
}

int bic_wait_stream(bic_ctx *c, void *producer_stream) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    cudaStream_t ps = (cudaStream_t)producer_stream;   // NULL = the legacy default stream
    if (ps == c->stream) return BIC_OK;
    if (!c->ev_wait) CU(cudaEventCreateWithFlags(&c->ev_wait, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_wait, ps));
    CU(cudaStreamWaitEvent(c->stream, c->ev_wait, 0));
    return BIC_OK;
}

int bic_dataset_fingerprint(bic_ctx *c, uint64_t *out) {
    if (!c || !out) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->data) return fail(c, BIC_ERR_NO_DATASET, "no dataset: call bic_set_dataset first");
    CU(cudaSetDevice(c->device));
    CU(c->fp_buf.ensure(sizeof(u64)));
    CU(cudaMemsetAsync(c->fp_buf.p, 0, sizeof(u64), c->stream));
    dim3 grid((unsigned)std::min<long long>(1024, (c->N / 16 + 255) / 256 + 1), (unsigned)c->n);
    k_fingerprint<<<grid, 256, 0, c->stream>>>(c->data, c->N, c->stride, c->n, c->fp_buf.as<u64>()); LAUNCH(c);
    CU(cudaGetLastError());
    u64 v = 0;
    CU(cudaMemcpyAsync(&v, c->fp_buf.p, sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    *out = v ^ mix_host((u64)c->N * 0x9E3779B97F4A7C15ULL + (u64)c->n);
    return BIC_OK;
}

int bic_set_iss(bic_ctx *c, double iss) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!(iss > 0.0)) return fail(c, BIC_ERR_ARG, "iss must be positive");
    CU(cudaSetDevice(c->device));
    if (iss != c->iss && c->cache_mode == 1) TRY(cache_clear(c));
    c->iss = iss;
    return BIC_OK;
}

int bic_sync(bic_ctx *c) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return BIC_OK;
}

int bic_set_dataset(bic_ctx *c, const uint8_t *codes, int64_t N, int32_t n, int64_t stride, const int32_t *card,
                    int is_device) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!codes || !card) return fail(c, BIC_ERR_ARG, "codes/card is NULL");
    if (n < 1 || n > NMAX) return fail(c, BIC_ERR_ARG, "n must be in 1..1024");
    if (N < 1 || N >= (1ll << 31)) return fail(c, BIC_ERR_ARG, "N must be in 1..2^31-1 (int32 count tables)");
    if (stride < N) return fail(c, BIC_ERR_ARG, "stride < N");
    for (int v = 0; v < n; ++v)
        if (card[v] < 1 || card[v] > 255) return fail(c, BIC_ERR_ARG, "cardinalities must be in 1..255");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    // the key width depends on n: drop the old cache entirely
    cache_free(c);
    c->lookups = c->misses = 0;
    if (c->data) { cudaFree(c->data); c->data = nullptr; }
    if (c->data2) { cudaFree(c->data2); c->data2 = nullptr; }
    if (c->d_card) { cudaFree(c->d_card); c->d_card = nullptr; }
    long long pstride = (N + 511) / 512 * 512;   // every column (and its 2-bit copy) 128-byte aligned; tail rows hold state 0
    CU(cudaMalloc(&c->data, (size_t)pstride * n));
    CU(cudaMalloc(&c->d_card, (size_t)n * sizeof(int)));
    CU(cudaMemsetAsync(c->data, 0, (size_t)pstride * n, c->stream));
    CU(cudaMemcpy2DAsync(c->data, (size_t)pstride, codes, (size_t)stride, (size_t)N, (size_t)n,
                         is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_card, card, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    c->N = N; c->n = n; c->stride = pstride; c->W64 = (n + 63) / 64; c->Wk = c->W64 + 1;
    c->card.assign(card, card + n);
    c->ntotal_dirty = true;
    // validate codes < card on the device
    TRY(header_reset(c));
    dim3 grid((unsigned)std::min<long long>(1024, (N / 16 + 255) / 256 + 1), (unsigned)n);
    k_validate<<<grid, 256, 0, c->stream>>>(c->data, N, pstride, n, c->d_card, &c->d_hdr->err); LAUNCH(c);
    CU(cudaGetLastError());
    TRY(header_fetch(c));
    if (c->h_hdr->err) {
        cudaFree(c->data);
        c->data = nullptr;
        c->N = 0;
        c->n = 0;
        return fail(c, BIC_ERR_BAD_CODE, "dataset holds a state code >= its declared cardinality");
    }
    // 2-bit shadow copy, worth it only when rows are streamed from HBM/L2 many times
    bool any_small = false;
    for (int v = 0; v < n; ++v) any_small = any_small || card[v] <= 4;
    // measured: at 100 k rows the packed path is 1.7x SLOWER (a thread runs only ~6 iterations of a
    // long unrolled body; per-item overhead and instruction fetch dominate), at 10 M rows 1.4x faster
    c->tune.from_env();   // tests switch the packed path per dataset
    c->all_packed = false;
    if (any_small && c->tune.pack2 && N >= c->tune.pack2_min_rows) {
        c->all_packed = true;
        for (int v = 0; v < n; ++v) c->all_packed = c->all_packed && card[v] <= 4;
        c->stride2 = pstride / 4;
        CU(cudaMalloc(&c->data2, (size_t)c->stride2 * n));
        CU(cudaMemsetAsync(c->data2, 0, (size_t)c->stride2 * n, c->stream));
        dim3 g2((unsigned)std::min<long long>(2048, (pstride / 16 + 255) / 256), (unsigned)n);
        k_pack2<<<g2, 256, 0, c->stream>>>(c->data, pstride, n, c->d_card, c->data2, c->stride2); LAUNCH(c);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));
    }
    return BIC_OK;
}

int bic_count_families(bic_ctx *c, const int32_t *node, const int64_t *parent_off, const int32_t *parents, int64_t F,
                       const int64_t *counts_off, int32_t *counts_out, int flags) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    TRY(begin_call(c, 0, false));
    if (F < 0 || (F > 0 && (!node || !parent_off || !counts_off || !counts_out))) return fail(c, BIC_ERR_ARG, "null pointer or negative F");
    if (F == 0) return BIC_OK;
    if (F > (1ll << 22)) return fail(c, BIC_ERR_ARG, "bic_count_families: at most 2^22 families per call");
    const bool dev = (flags & BIC_FLAG_DEVICE_PTRS) != 0;
    long long E = 0, cells_total = 0;
    if (!dev) { E = parent_off[F]; cells_total = counts_off[F]; }
    else {
        CU(cudaMemcpyAsync(&E, parent_off + F, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&cells_total, counts_off + F, sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    }
    const int *d_node = nullptr, *d_par = nullptr;
    const long long *d_off = nullptr, *d_coff = nullptr;
    TRY(stage_in(c, (const int *)node, (size_t)F, flags, c->in_stage, &d_node));
    TRY(stage_in(c, (const long long *)parent_off, (size_t)F + 1, flags, c->in_stage2, &d_off));
    TRY(stage_in(c, (const int *)parents, (size_t)std::max<long long>(E, 0), flags, c->in_stage3, &d_par));
    TRY(stage_in(c, (const long long *)counts_off, (size_t)F + 1, flags, c->in_stage4, &d_coff));
    CU(c->keybuf.ensure((size_t)F * c->Wk * sizeof(u64)));
    CU(c->cells_arr.ensure((size_t)F * sizeof(u32)));
    CU(c->class_jobs.ensure((size_t)F * NCLASS * sizeof(int)));
    CU(c->tmp_ll.ensure((size_t)F * 2 * sizeof(double)));
    TRY(header_reset(c));
    k_keys_csr<<<nblk(F, 256), 256, 0, c->stream>>>(d_off, d_par, d_node, F, c->n, c->W64, c->keybuf.as<u64>(), nullptr, c->d_hdr); LAUNCH(c);
    k_describe_direct<<<nblk(F, 256), 256, 0, c->stream>>>(c->keybuf.as<u64>(), c->W64, F, c->d_card, c->N, (u32)F, c->d_hdr,
                                                           c->cells_arr.as<u32>(), c->class_jobs.as<int>()); LAUNCH(c);
    CU(cudaGetLastError());
    TRY(header_fetch(c));
    TRY(check_header_err(c));
    TRY(run_count(c, c->keybuf.as<u64>(), 0, F, F, RunOut{c->tmp_ll.as<double>(), c->tmp_ll.as<double>() + F, 0}, true, false, false));
    int *d_out = counts_out;
    if (!dev) {
        CU(c->out_stage.ensure((size_t)std::max<long long>(cells_total, 1) * sizeof(int)));
        d_out = c->out_stage.as<int>();
    }
    k_copy_tables<<<(unsigned)F, 256, 0, c->stream>>>(c->arena.as<u32>(), c->table_off.as<u64>(), c->cells_arr.as<u32>(), d_coff, d_out, c->d_hdr); LAUNCH(c);
    CU(cudaGetLastError());
    TRY(header_fetch(c));
    TRY(check_header_err(c));
    if (!dev && cells_total) CU(cudaMemcpyAsync(counts_out, d_out, (size_t)cells_total * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return finish_call(c);
}

int bic_score_families(bic_ctx *c, const int32_t *node, const int64_t *parent_off, const int32_t *parents, int64_t F,
                       int metric, double *out, int flags) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    TRY(begin_call(c, metric, true));
    if (F < 0 || (F > 0 && (!node || !parent_off || !out))) return fail(c, BIC_ERR_ARG, "null pointer or negative F");
    if (flags & BIC_FLAG_NO_CACHE) TRY(cache_clear(c));
    const bool dev = (flags & BIC_FLAG_DEVICE_PTRS) != 0;
    const double pen = metric_penalty(c, metric);
    const long long Fs = 1ll << 22;
    std::vector<long long> h_off;
    for (long long f0 = 0; f0 < F; f0 += Fs) {
        long long Fc = std::min<long long>(Fs, F - f0);
        const int *d_node = nullptr, *d_par = nullptr;
        const long long *d_off = nullptr;
        if (dev) {
            d_node = (const int *)node + f0; d_off = (const long long *)parent_off + f0; d_par = (const int *)parents;
        } else {
            const long long *off = (const long long *)parent_off + f0;
            long long e0 = off[0], e1 = off[Fc];
            if (e1 < e0) return fail(c, BIC_ERR_ARG, "parent_off is not monotone");
            h_off.resize((size_t)Fc + 1);
            for (long long i = 0; i <= Fc; ++i) h_off[(size_t)i] = off[i] - e0;
            TRY(stage_in(c, (const int *)node + f0, (size_t)Fc, 0, c->in_stage, &d_node));
            TRY(stage_in(c, h_off.data(), (size_t)Fc + 1, 0, c->in_stage2, &d_off));
            TRY(stage_in(c, (const int *)parents + e0, (size_t)(e1 - e0), 0, c->in_stage3, &d_par));
            CU(cudaStreamSynchronize(c->stream));
        }
        CU(c->keybuf.ensure((size_t)Fc * c->Wk * sizeof(u64)));
        TRY(header_reset(c));
        k_keys_csr<<<nblk(Fc, 256), 256, 0, c->stream>>>(d_off, d_par, d_node, Fc, c->n, c->W64, c->keybuf.as<u64>(), nullptr, c->d_hdr); LAUNCH(c);
        CU(cudaGetLastError());
        TRY(resolve_instances(c, Fc, 0, (flags & BIC_FLAG_NO_DERIVE) != 0));
        double *d_out = out + f0;
        if (!dev) {
            CU(c->out_stage.ensure((size_t)Fc * sizeof(double)));
            d_out = c->out_stage.as<double>();
        }
        k_gather_fams<<<nblk(Fc, 256), 256, 0, c->stream>>>(c->inst.as<int>(), c->table, Fc, c->reg_ll, c->reg_np, pen, d_out); LAUNCH(c);
        CU(cudaGetLastError());
        if (!dev) CU(cudaMemcpyAsync(out + f0, d_out, (size_t)Fc * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        TRY(finish_call(c));
    }
    return BIC_OK;
}

int bic_score_dags_adj(bic_ctx *c, const uint8_t *adj, int64_t B, int metric, double *out, int64_t *n_invalid, int flags) {
    return score_dags(c, FMT_ADJ, adj, nullptr, B, metric, out, n_invalid, flags);
}

int bic_score_dags_csr(bic_ctx *c, const int64_t *off, const int32_t *parents, int64_t B, int metric, double *out,
                       int64_t *n_invalid, int flags) {
    return score_dags(c, FMT_CSR, off, parents, B, metric, out, n_invalid, flags);
}

int bic_score_dags_wire(bic_ctx *c, const uint8_t *labels, const uint32_t *ebits, int64_t B, int metric, double *out,
                        int64_t *n_invalid, int flags) {
    if (c && B > 0 && !ebits) return fail(c, BIC_ERR_ARG, "ebits is NULL");
    return score_dags(c, FMT_WIRE, labels, ebits, B, metric, out, n_invalid, flags);
}

int bic_score_dags_wire16(bic_ctx *c, const uint16_t *labels, const uint32_t *ebits, int32_t ewords, int64_t B, int metric,
                          double *out, int64_t *n_invalid, int flags) {
    if (c && B > 0 && !ebits) return fail(c, BIC_ERR_ARG, "ebits is NULL");
    if (c && ewords < 1) return fail(c, BIC_ERR_ARG, "ewords must be >= 1");
    return score_dags(c, FMT_WIRE16, labels, ebits, B, metric, out, n_invalid, flags, ewords);
}

int bic_cache_clear(bic_ctx *c) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    return cache_clear(c);
}

int bic_cache_reserve(bic_ctx *c, int64_t families) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->data) return fail(c, BIC_ERR_NO_DATASET, "no dataset: call bic_set_dataset first");
    CU(cudaSetDevice(c->device));
    long long extra = families - c->reg_count;
    return extra > 0 ? cache_ensure(c, extra) : BIC_OK;
}

int bic_cache_stats(bic_ctx *c, bic_cache_stats_t *out) {
    if (!c || !out) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    out->families = c->reg_count;
    out->capacity = c->reg_cap;
    out->lookups = c->lookups;
    out->misses = c->misses;
    out->bytes = (int64_t)(c->table_cap * sizeof(u32) + (size_t)c->reg_cap * (c->Wk * sizeof(u64) + 2 * sizeof(double)));
    return BIC_OK;
}

int bic_cache_export(bic_ctx *c, uint64_t *keys, double *terms, double *nparams, int64_t capacity, int64_t *families,
                     int32_t *kind) {
    if (!c || !families) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    *families = c->reg_count;
    if (kind) *kind = c->cache_mode;
    if (capacity == 0) return BIC_OK;
    if (capacity < c->reg_count || !keys || !terms || !nparams) return fail(c, BIC_ERR_ARG, "export buffers too small or NULL");
    if (!c->reg_count) return BIC_OK;
    CU(cudaSetDevice(c->device));
    CU(cudaMemcpyAsync(keys, c->regkeys, (size_t)c->reg_count * c->Wk * sizeof(u64), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(terms, c->reg_ll, (size_t)c->reg_count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(nparams, c->reg_np, (size_t)c->reg_count * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BIC_OK;
}

int bic_cache_import(bic_ctx *c, const uint64_t *keys, const double *terms, const double *nparams, int64_t families,
                     int32_t kind) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!c->data) return fail(c, BIC_ERR_NO_DATASET, "no dataset: call bic_set_dataset first");
    if (families < 0 || kind < 0 || kind > 2 || (families > 0 && (!keys || !terms || !nparams)))
        return fail(c, BIC_ERR_ARG, "bad import arguments");
    CU(cudaSetDevice(c->device));
    TRY(cache_clear(c));
    c->cache_mode = kind;
    if (!families) return BIC_OK;
    TRY(cache_ensure(c, families));
    CU(cudaMemcpyAsync(c->regkeys, keys, (size_t)families * c->Wk * sizeof(u64), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->reg_ll, terms, (size_t)families * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->reg_np, nparams, (size_t)families * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    // keys are validated on the device: node and parent bits inside the dataset, no self parent, no key twice
    TRY(header_reset(c));
    k_check_keys<<<nblk(families, 256), 256, 0, c->stream>>>(c->regkeys, c->W64, families, c->n, &c->d_hdr->err); LAUNCH(c);
    CU(cudaGetLastError());
    TRY(header_fetch(c));
    if (c->h_hdr->err) { cache_clear(c); return fail(c, BIC_ERR_ARG, "imported key names a node or parent outside the dataset, or a node as its own parent"); }
    k_rehash<<<nblk(families, 256), 256, 0, c->stream>>>(c->regkeys, c->Wk, families, c->table, (u32)(c->table_cap - 1)); LAUNCH(c);
    k_check_duplicates<<<nblk(families, 256), 256, 0, c->stream>>>(c->regkeys, c->Wk, families, c->table, (u32)(c->table_cap - 1), &c->d_hdr->err); LAUNCH(c);
    CU(cudaGetLastError());
    TRY(header_fetch(c));
    if (c->h_hdr->err) { cache_clear(c); return fail(c, BIC_ERR_ARG, "imported keys hold the same family twice"); }
    c->reg_count = families;
    return BIC_OK;
}

int bic_profile_enable(bic_ctx *c, int on) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    c->prof_on = on != 0;
    return BIC_OK;
}

int bic_profile_reset(bic_ctx *c) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    c->prof = bic_profile_t{};
    return BIC_OK;
}

int bic_profile_get(bic_ctx *c, bic_profile_t *out) {
    if (!c || !out) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    *out = c->prof;
    return BIC_OK;
}

int bic_comm_unique_id(uint8_t id_out[128]) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!id_out) return BIC_ERR_ARG;
    if (!g_nccl.load()) { g_create_error = g_nccl.err; return BIC_ERR_NCCL; }
    nccl_uid id;
    int rc = g_nccl.GetUniqueId(&id);
    if (rc != 0) { g_create_error = "ncclGetUniqueId failed"; return BIC_ERR_NCCL; }
    memcpy(id_out, id.internal, 128);
    return BIC_OK;
}

int bic_comm_init(bic_ctx *c, const uint8_t id[128], int rank, int world) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(c, BIC_ERR_ARG, "bad rank/world/id");
    {
        std::lock_guard<std::mutex> lk2(g_nccl_mu);
        if (!g_nccl.load()) return fail(c, BIC_ERR_NCCL, g_nccl.err);
    }
    CU(cudaSetDevice(c->device));
    xchg_release(c);
    if (c->comm) { g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
    nccl_uid uid;
    memcpy(uid.internal, id, 128);
    int rc = g_nccl.CommInitRank(&c->comm, world, uid, rank);
    if (rc != 0) {
        c->comm = nullptr;
        return fail(c, BIC_ERR_NCCL, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    }
    c->rank_id = rank;
    c->world = world;
    c->ntotal_dirty = true;
    TRY(cache_clear(c));   // cached terms were computed on this rank's rows only
    return BIC_OK;
}

int bic_comm_mode(bic_ctx *c, int mode) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    if (mode != BIC_SHARD_ROWS && mode != BIC_SHARD_FAMILIES) return fail(c, BIC_ERR_ARG, "unknown sharding mode");
    CU(cudaSetDevice(c->device));
    c->comm_mode = mode;
    c->ntotal_dirty = true;
    TRY(cache_clear(c));
    return BIC_OK;
}

int bic_comm_destroy(bic_ctx *c) {
    if (!c) return BIC_ERR_ARG;
    std::lock_guard<std::mutex> lk(c->mu);
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    xchg_release(c);
    if (c->comm) { g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
    c->world = 1; c->rank_id = 0;
    c->ntotal_dirty = true;
    TRY(cache_clear(c));
    return BIC_OK;
}

}  // extern "C"
