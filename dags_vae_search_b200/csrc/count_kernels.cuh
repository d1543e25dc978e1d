// k2 (family count) + k3 (fused fp64 log-likelihood reduce).
//
// Replaces the arithmetic behind bnlearn::score(net, data, type = "bic") (call site
// bnlearn_score.R:38): per family, mixed-radix parent configuration index per row,
// contingency counts N_ijk, then sum N_ijk ln(N_ijk / N_ij).
//
// Work item = (family job, row slice).  A CTA streams its slice of the k+1 state columns with
// 128-bit loads (16 rows per load), builds the cell index of each row in registers and counts
// into a histogram that is private to the CTA: in shared memory for classes 0..2, straight in
// HBM (L2 atomics) for class 3.  With one slice per family the fp64 reduce runs as the epilogue
// on the shared-memory table; otherwise slices merge into the HBM table and the last slice to
// finish (atomic ticket) reduces it.
#pragma once
#include "common.cuh"

namespace bic {

struct FamMetaC;

struct CountArgs {
    const uint8_t *data;     // [n][stride] uint8 state codes
    const uint8_t *data2;    // [n][stride2] 2-bit packed copy of the columns with <= 4 states (nullable)
    long long N;             // rows on this GPU
    long long stride;
    long long stride2;
    const int *card;
    int W64;
    const u64 *keys;         // family keys; job j uses keys + (key_base + j) * (W64 + 1)
    long long key_base;
    const int *jobs;         // job ids of this launch (one class)
    int njobs;               // how many
    int S;                   // row slices per family
    int P;                   // RANGE kernel: cell sub-range passes per (family, slice) at most
    u32 span;                // RANGE kernel: cells of one sub-range (the shared-memory table of a CTA)
    u32 cap_words;           // shared-memory words available to one CTA's table(s)
    u32 stage_words;         // dynamic shared memory of the launch in words (staging buffer of the HBM-table reduce)
    u32 *arena;              // HBM count tables
    const u32 *need;         // per job: cells if its table lives in HBM, else 0 (nullable)
    const u64 *table_off;    // per job: offset into arena (valid where need != 0)
    u32 *done;               // per job: slice tickets
    long long out_base;      // job j writes ll_out / np_out [out_base + j] (= key_base for the registry, 0 for staged terms)
    double *ll_out;
    double *np_out;          // (r - 1) * q
    int reduce;              // 0: count only (row-sharded: reduce after the all-reduce)
    const int *donor;        // per job: job whose table this one is marginalised from, or -1 (nullable)
    u32 *cellbuf;            // RANGE kernel: cell index of every row, [position in the class-3 job list][stride] (k_cells), or NULL
    int swizzle;             // packed path, un-replicated tables: bank swizzle of the cell index (swz_off)
    u32 tier_lo, tier_hi;    // the launch counts only the families with tier_lo <= cells <= tier_hi (0 / 0: all): one class list,
                             //   two launches with different CTA shapes (run_count)
    const int2 *items3;      // RANGE kernel: (job, pass) per work item of a row slice (k_range_items), or NULL: njobs * P items
    int nitems3;
    int c3_u16;              // RANGE kernel: 16-bit counters, two per word (sub-ranges of 2 * span cells: half the passes), spilled
                             //   into the HBM table between barrier-separated phases so that none can overflow (count_rows_r16)
    int topsplit;            // RANGE kernel: sub-ranges along the first parent's states where range_plan() allows it
    int u8_narrow;           // uint8 path of classes 0 / 1: 8-byte loads (experiment)
    int p2_two;              // packed path, families of <= 3 columns: two 64-row groups in flight per thread (datasets beyond L2)
    int u8_two;              // uint8 path, families of <= 4 columns: row groups in flight per thread (0: one, 2: two, 3: up to four for k <= 1)
    int tma;                 // uint8 path of classes 0 / 1: rows staged through a shared-memory ring with bulk copies (experiment)
    const FamMetaC *meta;    // per job: the decoded family (k_decode_jobs), or NULL: thread 0 of every CTA decodes the key
    u32 k30, k28, k26;       // 2^30, 2^28, 2^26 (opaque to the compiler; -DBIC_UNPACK_FMA experiment)
    int p2_vec;              // packed path: 32-bit words of a column one thread loads per iteration (4, 2 or 1)
    int bd_mode;             // 0: log-likelihood terms; 1: BDeu with imaginary sample size iss; 2: K2
    double iss;
    // Row-sharded runs, fused count + reduce-scatter (push != 0): the CTA that completes a family's
    // local table stores it into slot `rank` of the owner rank's exchange buffer over NVLink
    // (peer[owner[j]] + rank * xcap + xoff[j]); the owner sums the `world` slots in k_reduce_tables.
    int push;
    int rank, world;
    const int *owner;        // per job: rank that reduces (and derives from) the family's global table
    const u64 *xoff;         // per job: cell offset inside a slot of the owner's exchange buffer
    u32 *const *peer;        // [world] exchange buffers (peer-mapped device pointers; own buffer at [rank])
    u64 xcap;                // cells per slot
    int writeback;           // k_reduce_tables: also store the summed table into the local arena (donors of derived families)
};

struct FamMeta {
    int k, node, r;
    u32 q, cells;
    int small;  // every column of the family has <= 4 states (2-bit packed copy usable)
    u32 R;      // lane replicas of the shared-memory table (power of two, <= 32)
    u32 mul;    // byte offset of a cell = cell * mul  (mul = 4 * R)
    u32 lo4, span4;   // RANGE kernel: byte offset of the CTA's first cell, bytes of its sub-range
    int par[KMAX];
    u32 rad[KMAX];
};

// The decoded family as k_decode_jobs parks it per job (128 bytes): a count CTA fetches it with one
// coalesced load instead of having thread 0 walk the key and the cardinalities while 255 threads
// wait at the first barrier (ncu source page, version k: 11 % of all warp samples sat there).
struct __align__(16) FamMetaC {
    int k, node, r;
    u32 q, cells;
    int small;
    uint16_t par[KMAX];
    uint8_t rad[KMAX];
    u32 pad[2];
};
static_assert(sizeof(FamMetaC) == 128, "one 128-byte record per job");

// Thread 0 decodes the key.  Parents ascending, first parent most significant; parents with a
// single state contribute nothing to the index and are dropped.
__device__ __forceinline__ void decode_family(const u64 *key, int W64, const int *__restrict__ card, FamMeta &m) {
    m.node = (int)key[0];
    m.r = card[m.node];
    int k = 0;
    int small = m.r <= 4;
    u32 q = 1;
    for (int w = 0; w < W64; ++w) {
        u64 bits = key[1 + w];
        while (bits) {
            int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            int p = w * 64 + b;
            int c = card[p];
            if (c > 1 && k < KMAX) {
                m.par[k] = p;
                m.rad[k] = (u32)c;
                q *= (u32)c;
                small = small && c <= 4;
                ++k;
            }
        }
    }
    m.k = k;
    m.q = q;
    m.cells = q * (u32)m.r;
    m.small = small;
}

__global__ void k_decode_jobs(const u64 *__restrict__ keys, long long key_base, int W64, const int *__restrict__ card, long long njobs,
                              FamMetaC *out) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    FamMeta m;
    decode_family(keys + (key_base + j) * (long long)(W64 + 1), W64, card, m);
    FamMetaC c;
    c.k = m.k; c.node = m.node; c.r = m.r; c.q = m.q; c.cells = m.cells; c.small = m.small;
    for (int a = 0; a < KMAX; ++a) {
        c.par[a] = a < m.k ? (uint16_t)m.par[a] : 0;
        c.rad[a] = a < m.k ? (uint8_t)m.rad[a] : 0;
    }
    c.pad[0] = c.pad[1] = 0;
    out[j] = c;
}

// Work items of the class-3 sub-range kernel.  The grid used to hold njobs x P items per row slice, P the
// pass count of the largest table of the launch; families with smaller tables left their surplus items
// empty, and with one CTA per SM an empty item is an idle SM (ncu on the diabetes-shaped step: SMs active
// half of the launch).  This lists (job, pass) for exactly the passes every family needs; the host sizes the
// row slices for that number of items.  One block; `cnt` = class-3 jobs of the launch.
__global__ void __launch_bounds__(1024) k_range_items(const int *__restrict__ jobs, int cnt, const FamMetaC *__restrict__ meta, u32 span,
                                                      int u16, int topsplit, int2 *items, u32 cap) {
    __shared__ u32 s_scan[1024];
    __shared__ u32 s_carry;
    const int tid = (int)threadIdx.x;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < cnt; base += 1024) {
        const int i = base + tid;
        u32 p = 0;
        int job = -1;
        if (i < cnt) {
            job = jobs[i];
            const int kk = meta[job].k;
            const u32 rad0 = kk > 0 ? (u32)meta[job].rad[0] : 1u;
            p = (u16 && kk <= 6) ? range_plan(meta[job].cells, kk, rad0, 2u * span, false).passes
                                 : range_plan(meta[job].cells, kk, rad0, span, topsplit != 0).passes;
        }
        s_scan[tid] = p;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const u32 x = tid >= o ? s_scan[tid - o] : 0u;
            __syncthreads();
            s_scan[tid] += x;
            __syncthreads();
        }
        const u32 first = s_carry + s_scan[tid] - p;
        for (u32 q = 0; q < p; ++q)
            if (first + q < cap) items[first + q] = make_int2(job, (int)q);
        __syncthreads();
        if (tid == 1023) s_carry += s_scan[1023];
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// Cell index arithmetic.  One 128-bit load holds 16 consecutive rows of one column.  The
// mixed-radix index is built on packed lanes so that one IMAD (FMA pipe) advances several rows
// at once and the per-row byte extraction (ALU pipe, the measured bottleneck of the first
// version: 77 % ALU, 60 % DRAM) happens once per row instead of once per row per column:
//   MODE_U8  (cells <= 64):        8-bit lanes, 4 rows per register, no unpack at all;
//   MODE_U16 (cells*R <= 16383): 16-bit lanes, 2 rows per register, bytes unpacked with PRMT,
//                                 index pre-multiplied by `mul` so the extracted lane is the
//                                 byte offset that feeds the shared-memory atomic directly;
//   MODE_U32 (anything else):     one register per row.
//
// Lane replicas.  After the arithmetic moved to packed lanes the limiter became the shared-
// memory data pipe: 2.2 wavefronts per warp atomic, 54 % of them bank conflicts between
// different cells (ncu, profiles/r01b).  Small tables are therefore kept in R bank-interleaved
// replicas, slot = cell * R + (lane % R): with R = 32 lane l only ever touches bank l, so a
// warp atomic is one wavefront whatever the data; R = 16/8/4/2 cut conflicts proportionally.
// Replicas are summed once, after the row loop.
enum { MODE_U8 = 0, MODE_U16 = 1, MODE_U32 = 2 };
constexpr int REPL_MAX_PER_THREAD = 16;   // replicated tables hold at most 16 cells per thread (compact_replicas)

__device__ __forceinline__ int count_mode(u32 cells, u32 R) {
    return cells <= 64u ? MODE_U8 : cells * R <= 16383u ? MODE_U16 : MODE_U32;
}

template <bool GLOBAL>
__device__ __forceinline__ void bump_off(u32 *hist, u32 byte_off) {
    atomicAdd(reinterpret_cast<u32 *>(reinterpret_cast<char *>(hist) + byte_off), 1u);   // ATOMS.POPC.INC / RED
}

// Bank swizzle of un-replicated tables: slot = cell ^ ((cell >> 5) & 31), applied to the byte offset.
// A warp's 32 increments then spread over the banks by ten index bits instead of five: the low
// digits alone (child and last parent, often dominated by one or two states) left class-1 tables at
// 4.1 wavefronts per warp atomic, 75 % of them bank conflicts (ncu).  The table is put back in
// order before it is merged or reduced (unswizzle_table).
__device__ __forceinline__ u32 swz_off(u32 off) { return off ^ ((off >> 5) & 0x7cu); }

// byte offsets of the 16 rows of a group, in row order; rows >= N are skipped.  RANGE: the CTA
// owns the cells [lo4, lo4 + span4) (byte offsets) of a table too large for shared memory and
// ignores the rows that fall outside (another pass counts them).
template <bool GLOBAL, bool RANGE = false>
__device__ __forceinline__ void bump16(u32 *hist, const u32 (&off)[16], long long row0, long long N, u32 lo4 = 0,
                                       u32 span4 = 0) {
    if (RANGE) {
        const int nv = row0 + 16 <= N ? 16 : (int)(N - row0);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const u32 d = off[b] - lo4;
            if (b < nv && d < span4) bump_off<false>(hist, d);
        }
    } else if (row0 + 16 <= N) {
#pragma unroll
        for (int b = 0; b < 16; ++b) bump_off<GLOBAL>(hist, off[b]);
    } else {
        int nv = (int)(N - row0);
#pragma unroll
        for (int b = 0; b < 16; ++b)
            if (b < nv) bump_off<GLOBAL>(hist, off[b]);
    }
}

template <int K>
__device__ __forceinline__ void cells_u8(const uint4 (&w)[K + 1], const u32 (&rad)[K + 1], u32 mul, u32 (&off)[16]) {
    u32 acc[4] = {w[0].x, w[0].y, w[0].z, w[0].w};
#pragma unroll
    for (int a = 1; a <= K; ++a) {
        const u32 ws[4] = {w[a].x, w[a].y, w[a].z, w[a].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = acc[i] * rad[a] + ws[i];   // 4 rows per IMAD, lanes stay < 64
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int b = 0; b < 4; ++b) off[i * 4 + b] = __dp4a(acc[i], mul << (8 * b), 0u);   // byte b * mul in one IDP.4A (mul <= 128)
}

template <int K>
__device__ __forceinline__ void cells_u16(const uint4 (&w)[K + 1], const u32 (&rad)[K + 1], u32 mul, u32 (&off)[16]) {
    u32 acc[8];
#pragma unroll
    for (int a = 0; a <= K; ++a) {
        const u32 ws[4] = {w[a].x, w[a].y, w[a].z, w[a].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            u32 lo = __byte_perm(ws[i], 0u, 0x4140u);   // rows 4i, 4i+1 in 16-bit lanes
            u32 hi = __byte_perm(ws[i], 0u, 0x4342u);   // rows 4i+2, 4i+3
            if (a == 0) {
                acc[2 * i] = lo;
                acc[2 * i + 1] = hi;
            } else {
                acc[2 * i] = acc[2 * i] * rad[a] + lo;   // 2 rows per IMAD, lanes stay < 16384
                acc[2 * i + 1] = acc[2 * i + 1] * rad[a] + hi;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        u32 a4 = acc[i] * mul;                           // lanes <= 65532
        off[2 * i] = a4 & 0xffffu;
        off[2 * i + 1] = a4 >> 16;
    }
}

template <int K>
__device__ __forceinline__ void cells_u32(const uint4 (&w)[K + 1], const u32 (&rad)[K + 1], u32 mul, u32 (&off)[16]) {
#pragma unroll
    for (int b = 0; b < 16; ++b) off[b] = 0;
#pragma unroll
    for (int a = 0; a <= K; ++a) {
        const u32 ws[4] = {w[a].x, w[a].y, w[a].z, w[a].w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) off[i * 4 + b] = off[i * 4 + b] * rad[a] + ((ws[i] >> (8 * b)) & 0xffu);
    }
#pragma unroll
    for (int b = 0; b < 16; ++b) off[b] *= mul;
}

// K parents known at compile time: all K+1 column loads of a row group are issued before any
// is consumed (K+1 independent 16-byte loads in flight per thread).
template <int K, int MODE, bool GLOBAL, int THREADS, bool RANGE = false>
__device__ __forceinline__ void count_rows_k(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                             long long N, long long v0, long long v1, u32 *hist) {
    const uint8_t *cp[K + 1];
    u32 rad[K + 1];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    const u32 mul = m.mul;
    for (long long v = v0 + threadIdx.x; v < v1; v += THREADS) {
        uint4 w[K + 1];
#pragma unroll
        for (int a = 0; a <= K; ++a) w[a] = ld_stream_v4(cp[a] + v * 16);
        u32 off[16];
        if (MODE == MODE_U8) cells_u8<K>(w, rad, mul, off);
        else if (MODE == MODE_U16) cells_u16<K>(w, rad, mul, off);
        else cells_u32<K>(w, rad, mul, off);
        bump16<GLOBAL, RANGE>(hist, off, v * 16, N, m.lo4, m.span4);
    }
}

// The same with the loads of U row groups in flight per thread (families of <= 4 columns, so that
// U (k + 1) vector registers fit): on datasets that stream from HBM the uint8 path is bound by
// the bytes in flight, not by the data pipe (halving them with 8-byte loads cost 50 %; doubling
// them took the diabetes-shaped class-0 launch from 1.22 to 1.14 ms).
template <int K, int MODE, int THREADS, int U>
__device__ __forceinline__ void count_rows_ku(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                              long long N, long long v0, long long v1, u32 *hist) {
    const uint8_t *cp[K + 1];
    u32 rad[K + 1];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    const u32 mul = m.mul;
    for (long long v = v0 + threadIdx.x; v < v1; v += (long long)U * THREADS) {
        uint4 w[U][K + 1];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int a = 0; a <= K; ++a)
                w[u][a] = (u == 0 || v + (long long)u * THREADS < v1) ? ld_stream_v4(cp[a] + (v + (long long)u * THREADS) * 16) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long vu = v + (long long)u * THREADS;
            if (u == 0 || vu < v1) {
                u32 off[16];
                if (MODE == MODE_U8) cells_u8<K>(w[u], rad, mul, off);
                else if (MODE == MODE_U16) cells_u16<K>(w[u], rad, mul, off);
                else cells_u32<K>(w[u], rad, mul, off);
                bump16<false, false>(hist, off, vu * 16, N);
            }
        }
    }
}

template <int K, int MODE, int THREADS>
__device__ __forceinline__ void count_rows_k2(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                              long long N, long long v0, long long v1, u32 *hist, int two) {
    constexpr int UMAX = K == 0 ? 4 : K == 1 ? 3 : 2;
    if (two >= 3 && UMAX > 2) count_rows_ku<K, MODE, THREADS, UMAX>(m, data, stride, N, v0, v1, hist);
    else count_rows_ku<K, MODE, THREADS, 2>(m, data, stride, N, v0, v1, hist);
}

// Class 3, top split (range_plan in common.cuh): the CTA owns the states [s0, s0 + ns) of the first
// parent.  The index below that parent (< 16384 cells) is built on 16-bit packed lanes exactly as
// cells_u16 does for the smaller classes; the first parent costs one byte extraction, one compare
// and one IMAD per row, and rows of other states stop there.  hist = the CTA's sub-range, cell
// (s - s0) * low + rest at byte offset 4 * that.
template <int K, int THREADS>
__device__ __forceinline__ void count_rows_top(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                               long long N, long long v0, long long v1, u32 *hist, u32 s0, u32 ns, u32 low4) {
    static_assert(K >= 1, "the first parent is the split axis");
    const uint8_t *cp[K + 1];
    u32 rad[K + 1];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    for (long long v = v0 + threadIdx.x; v < v1; v += THREADS) {
        uint4 w[K + 1];
#pragma unroll
        for (int a = 0; a <= K; ++a) w[a] = ld_stream_v4(cp[a] + v * 16);
        u32 acc[8];
#pragma unroll
        for (int a = 1; a <= K; ++a) {
            const u32 ws[4] = {w[a].x, w[a].y, w[a].z, w[a].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                u32 lo = __byte_perm(ws[i], 0u, 0x4140u);   // rows 4i, 4i+1 in 16-bit lanes
                u32 hi = __byte_perm(ws[i], 0u, 0x4342u);   // rows 4i+2, 4i+3
                if (a == 1) {
                    acc[2 * i] = lo;
                    acc[2 * i + 1] = hi;
                } else {
                    acc[2 * i] = acc[2 * i] * rad[a] + lo;   // lanes stay < 16384
                    acc[2 * i + 1] = acc[2 * i + 1] * rad[a] + hi;
                }
            }
        }
        u32 off[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            u32 a4 = acc[i] * 4u;                            // lanes <= 65532
            off[2 * i] = a4 & 0xffffu;
            off[2 * i + 1] = a4 >> 16;
        }
        const u32 ts[4] = {w[0].x, w[0].y, w[0].z, w[0].w};
        const long long row0 = v * 16;
        const int nv = row0 + 16 <= N ? 16 : (int)(N - row0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const u32 d = __byte_perm(ts[i], 0u, 0x4440u + b) - s0;
                if (4 * i + b < nv && d < ns) bump_off<false>(hist, d * low4 + off[4 * i + b]);
            }
    }
}

// Class 3 with 16-bit counters (BIC_C3_U16): a sub-range holds 2 * H cells (H = the CTA's words), cell d
// of the sub-range in word d mod H, low half for d < H and high half above - neighbouring cells stay in
// neighbouring banks - so a table needs half the passes over the rows.  No counter can overflow: the
// row loop runs in phases of at most 48 K rows per CTA, and between two phases (block-wide barriers)
// every half that has reached 16 384 is added to the HBM table and cleared; 16 383 + 49 152 = 65 535.
template <int K, int THREADS>
__device__ __forceinline__ void count_rows_r16(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                               long long N, long long v0, long long v1, u32 *hist, u32 lo, u32 span,
                                               u32 H, u32 *__restrict__ tab) {
    constexpr int PH = 49152 / (16 * THREADS);   // 16-row groups per thread and phase
    static_assert(PH >= 1, "a phase adds at most 49152 rows");
    const uint8_t *cp[K + 1];
    u32 rad[K + 1];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    for (long long base = v0; base < v1; base += (long long)PH * THREADS) {
#pragma unroll 1
        for (int i = 0; i < PH; ++i) {
            const long long v = base + (long long)i * THREADS + threadIdx.x;
            if (v >= v1) break;
            uint4 w[K + 1];
#pragma unroll
            for (int a = 0; a <= K; ++a) w[a] = ld_stream_v4(cp[a] + v * 16);
            u32 idx[16];
            cells_u32<K>(w, rad, 1u, idx);
            const int nv = v * 16 + 16 <= N ? 16 : (int)(N - v * 16);
#pragma unroll
            for (int b = 0; b < 16; ++b) {
                const u32 d = idx[b] - lo;
                if (b < nv && d < span) {
                    const bool up = d >= H;
                    atomicAdd(hist + (up ? d - H : d), up ? 0x10000u : 1u);
                }
            }
        }
        if (base + (long long)PH * THREADS < v1) {   // uniform: another phase follows
            __syncthreads();
            for (u32 c = threadIdx.x * 4u; c < H; c += THREADS * 4u) {   // H is a multiple of 4
                uint4 q = *reinterpret_cast<const uint4 *>(hist + c);
                if ((q.x | q.y | q.z | q.w) & 0xC000C000u) {
                    u32 qs[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const u32 l16 = qs[e] & 0xffffu, h16 = qs[e] >> 16;
                        if (l16 >= 0x4000u) { atomicAdd(tab + lo + c + e, l16); qs[e] &= 0xffff0000u; }
                        if (h16 >= 0x4000u) { atomicAdd(tab + lo + H + c + e, h16); qs[e] &= 0x0000ffffu; }
                    }
                    *reinterpret_cast<uint4 *>(hist + c) = make_uint4(qs[0], qs[1], qs[2], qs[3]);
                }
            }
            __syncthreads();
        }
    }
}

template <int K, bool GLOBAL, int THREADS, bool RANGE = false>
__device__ __forceinline__ void count_rows_mode(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                                long long N, long long v0, long long v1, u32 *hist, int two = 0) {
    if (GLOBAL || RANGE) {   // class 3 tables are far above the packed-lane limits
        count_rows_k<K, MODE_U32, GLOBAL, THREADS, RANGE>(m, data, stride, N, v0, v1, hist);
        return;
    }
    if (K <= 3 && two) {
        switch (count_mode(m.cells, m.R)) {
            case MODE_U8: count_rows_k2<K <= 3 ? K : 0, MODE_U8, THREADS>(m, data, stride, N, v0, v1, hist, two); break;
            case MODE_U16: count_rows_k2<K <= 3 ? K : 0, MODE_U16, THREADS>(m, data, stride, N, v0, v1, hist, two); break;
            default: count_rows_k2<K <= 3 ? K : 0, MODE_U32, THREADS>(m, data, stride, N, v0, v1, hist, two); break;
        }
        return;
    }
    switch (count_mode(m.cells, m.R)) {
        case MODE_U8: count_rows_k<K, MODE_U8, GLOBAL, THREADS>(m, data, stride, N, v0, v1, hist); break;
        case MODE_U16: count_rows_k<K, MODE_U16, GLOBAL, THREADS>(m, data, stride, N, v0, v1, hist); break;
        default: count_rows_k<K, MODE_U32, GLOBAL, THREADS>(m, data, stride, N, v0, v1, hist); break;
    }
}

// ---------------------------------------------------------------------------------------
// TMA-staged row tiles (experiment knob BIC_TMA=1, uint8 path of classes 0 and 1).  One elected
// thread streams the k+1 column segments of a tile into a shared-memory ring with bulk copies
// (cp.async.bulk, SASS UBLKCP) that complete on an mbarrier; all threads read their 8 rows per
// column back with LDS.64.  The copies land in shared memory without passing through L1, so the
// in-flight rows no longer compete with the lane-replicated tables for the SM's unified
// L1 / shared memory; the read-back costs the same data-pipe wavefronts as the load return it
// replaces (see DESIGN.md for the measured outcome).
constexpr int TMA_STAGES = 2;
constexpr int TMA_NW = 2;                        // 32-bit words (4 rows each) per thread per column per tile
constexpr int TMA_MAXCOLS = 7;
__host__ __device__ constexpr u32 tma_ring_bytes(int threads) { return (u32)(TMA_STAGES * TMA_MAXCOLS * threads * 4 * TMA_NW); }

__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    asm volatile(
        "{\n .reg .pred p;\n TMA_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra TMA_DONE;\n bra TMA_WAIT;\n "
        "TMA_DONE:\n}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// cell byte offsets of 4 * NW consecutive rows from NW words per column (the word-count-generic
// forms of cells_u8 / cells_u16 / cells_u32 above)
template <int K, int NW, int MODE>
__device__ __forceinline__ void cells_words(const u32 (&w)[K + 1][NW], const u32 (&rad)[K + 1], u32 mul, u32 (&off)[4 * NW]) {
    if (MODE == MODE_U8) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            u32 acc = w[0][i];
#pragma unroll
            for (int a = 1; a <= K; ++a) acc = acc * rad[a] + w[a][i];
#pragma unroll
            for (int b = 0; b < 4; ++b) off[i * 4 + b] = __byte_perm(acc, 0u, 0x4440u + b) * mul;
        }
    } else if (MODE == MODE_U16) {
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            u32 lo = 0, hi = 0;
#pragma unroll
            for (int a = 0; a <= K; ++a) {
                const u32 l = __byte_perm(w[a][i], 0u, 0x4140u), h = __byte_perm(w[a][i], 0u, 0x4342u);
                lo = a == 0 ? l : lo * rad[a] + l;
                hi = a == 0 ? h : hi * rad[a] + h;
            }
            lo *= mul;
            hi *= mul;
            off[i * 4 + 0] = lo & 0xffffu; off[i * 4 + 1] = lo >> 16;
            off[i * 4 + 2] = hi & 0xffffu; off[i * 4 + 3] = hi >> 16;
        }
    } else {
#pragma unroll
        for (int i = 0; i < NW; ++i)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                u32 c = 0;
#pragma unroll
                for (int a = 0; a <= K; ++a) c = c * rad[a] + ((w[a][i] >> (8 * b)) & 0xffu);
                off[i * 4 + b] = c * mul;
            }
    }
}

// uint8 path with 8-byte loads (experiment knob BIC_U8_NARROW=1): half the bytes in flight per thread,
// twice the load instructions — affordable where the ALU pipe has headroom (the uint8 path; on the
// packed path narrower loads lost) and meant to make room for 48 KB class-0 tables there as well.
template <int K, int MODE, int THREADS>
__device__ __forceinline__ void count_rows_narrow(const FamMeta &m, const uint8_t *__restrict__ data, long long stride, long long N,
                                                  long long v0, long long v1, u32 *hist) {
    constexpr int C = K + 1;
    const uint8_t *cp[C];
    u32 rad[C];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    const u32 mul = m.mul;
    for (long long h = v0 * 2 + threadIdx.x; h < v1 * 2; h += THREADS) {   // 8-row half vectors
        u32 w[C][2];
#pragma unroll
        for (int a = 0; a < C; ++a)
            asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(w[a][0]), "=r"(w[a][1]) : "l"(cp[a] + h * 8));
        u32 off[8];
        cells_words<K, 2, MODE>(w, rad, mul, off);
        const long long row0 = h * 8;
        const int nv = row0 + 8 <= N ? 8 : (int)max(0ll, N - row0);
#pragma unroll
        for (int b = 0; b < 8; ++b)
            if (b < nv) bump_off<false>(hist, off[b]);
    }
}

template <int K, int THREADS>
__device__ __forceinline__ void count_rows_narrow_mode(const FamMeta &m, const uint8_t *__restrict__ data, long long stride, long long N,
                                                       long long v0, long long v1, u32 *hist) {
    switch (count_mode(m.cells, m.R)) {
        case MODE_U8: count_rows_narrow<K, MODE_U8, THREADS>(m, data, stride, N, v0, v1, hist); break;
        case MODE_U16: count_rows_narrow<K, MODE_U16, THREADS>(m, data, stride, N, v0, v1, hist); break;
        default: count_rows_narrow<K, MODE_U32, THREADS>(m, data, stride, N, v0, v1, hist); break;
    }
}

// [v0, v1): the CTA's slice in 16-row vectors; ring: TMA_STAGES x (K + 1) tiles of THREADS * 4 * NW bytes
template <int K, int MODE, int THREADS>
__device__ __forceinline__ void count_rows_tma(const FamMeta &m, const uint8_t *__restrict__ data, long long stride, long long N,
                                               long long v0, long long v1, u32 *hist, uint8_t *ring, u64 *full) {
    constexpr int NW = TMA_NW, C = K + 1;
    constexpr u32 TILE = THREADS * 4 * NW;                 // bytes (= rows) of one column segment
    const uint8_t *cp[C];
    u32 rad[C];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    const u32 mul = m.mul;
    const long long r0 = v0 * 16, r1 = v1 * 16;            // 16-byte granularity: r1 may exceed N inside the padded stride
    const long long ntiles = (r1 - r0 + TILE - 1) / TILE;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long t) {
        const int s = (int)(t % TMA_STAGES);
        const long long rb = r0 + t * TILE;
        const u32 bytes = (u32)min((long long)TILE, r1 - rb);
        mbar_expect_tx(&full[s], bytes * C);
#pragma unroll
        for (int a = 0; a < C; ++a) bulk_g2s(ring + ((size_t)s * C + a) * TILE, cp[a] + rb, bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (long long t = 0; t < TMA_STAGES - 1 && t < ntiles; ++t) issue(t);
    for (long long t = 0; t < ntiles; ++t) {
        const int s = (int)(t % TMA_STAGES);
        if (threadIdx.x == 0 && t + TMA_STAGES - 1 < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads of the stage are ordered before its refill
            issue(t + TMA_STAGES - 1);
        }
        mbar_wait(&full[s], (u32)((t / TMA_STAGES) & 1));
        const long long row = r0 + t * TILE + (long long)threadIdx.x * (4 * NW);
        if (row < N && row < r1) {   // the last tile of a slice is partly filled: what lies beyond r1 is another slice's
            u32 w[C][NW];
#pragma unroll
            for (int a = 0; a < C; ++a) {
                const uint2 x = *reinterpret_cast<const uint2 *>(ring + ((size_t)s * C + a) * TILE + threadIdx.x * (4 * NW));
                w[a][0] = x.x;
                w[a][1] = x.y;
            }
            u32 off[4 * NW];
            cells_words<K, NW, MODE>(w, rad, mul, off);
            const int nv = row + 4 * NW <= N ? 4 * NW : (int)(N - row);
#pragma unroll
            for (int b = 0; b < 4 * NW; ++b)
                if (b < nv) bump_off<false>(hist, off[b]);
        }
        __syncthreads();   // everyone is done with stage s before it is refilled
    }
}

template <int K, int THREADS>
__device__ __forceinline__ void count_rows_tma_mode(const FamMeta &m, const uint8_t *__restrict__ data, long long stride, long long N,
                                                    long long v0, long long v1, u32 *hist, uint8_t *ring, u64 *full) {
    switch (count_mode(m.cells, m.R)) {
        case MODE_U8: count_rows_tma<K, MODE_U8, THREADS>(m, data, stride, N, v0, v1, hist, ring, full); break;
        case MODE_U16: count_rows_tma<K, MODE_U16, THREADS>(m, data, stride, N, v0, v1, hist, ring, full); break;
        default: count_rows_tma<K, MODE_U32, THREADS>(m, data, stride, N, v0, v1, hist, ring, full); break;
    }
}

// ---------------------------------------------------------------------------------------
// 2-bit packed path (every column of the family has <= 4 states).  One 128-bit load now holds
// 64 rows of a column, so the L1TEX data pipe moves a quarter of the bytes.  A 32-bit word (16
// rows) is opened into 4 byte-lane registers with (W >> 2s) & 0x03030303: register s, byte lane
// B = row 4B + s.  The mixed-radix index is built in byte lanes (4 rows per IMAD): the last
// four columns (radices <= 4, so < 256 combinations) form the low group, the up to three
// columns before them the high group; both are widened to 16-bit lanes with PRMT and combined
// as hi * (P_low * mul) + lo * mul, which is the byte offset of the counter.
#ifdef BIC_UNPACK_FMA
// experiment: the three shifts as IMAD.HI by 2^30 / 2^28 / 2^26 (kernel parameters, opaque to the
// compiler) — FMA pipe instead of the ALU pipe that is at 80 %
__device__ __forceinline__ void unpack2(u32 W, u32 (&u)[4], u32 k30, u32 k28, u32 k26) {
    u[0] = W & 0x03030303u;
    u[1] = __umulhi(W, k30) & 0x03030303u;
    u[2] = __umulhi(W, k28) & 0x03030303u;
    u[3] = __umulhi(W, k26) & 0x03030303u;
}
#else
__device__ __forceinline__ void unpack2(u32 W, u32 (&u)[4], u32, u32, u32) {
    u[0] = W & 0x03030303u;
    u[1] = (W >> 2) & 0x03030303u;
    u[2] = (W >> 4) & 0x03030303u;
    u[3] = (W >> 6) & 0x03030303u;
}
#endif

// VEC = 32-bit words (16 rows each) of every column a thread loads per iteration: 4 (one 128-bit
// load = 64 rows), 2 or 1.  Fewer bytes in flight per thread leave more of the SM's unified
// L1 / shared memory to the lane-replicated tables (the rows come from L2, so short loads still
// cover the latency); the arithmetic per word is the same.
template <int VEC> struct P2Load;
template <> struct P2Load<4> {
    static __device__ __forceinline__ void ld(const uint8_t *p, u32 (&w)[4]) {
        uint4 r = ld_stream_v4(p);
        w[0] = r.x; w[1] = r.y; w[2] = r.z; w[3] = r.w;
    }
};
template <> struct P2Load<2> {
    static __device__ __forceinline__ void ld(const uint8_t *p, u32 (&w)[2]) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "l"(p));
    }
};
template <> struct P2Load<1> {
    static __device__ __forceinline__ void ld(const uint8_t *p, u32 (&w)[1]) {
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w[0]) : "l"(p));
    }
};

#ifndef BIC_P2_PAIRS
#define BIC_P2_PAIRS 1
#endif
#ifndef BIC_P2_DP4A
#define BIC_P2_DP4A 1
#endif
template <int K, int VEC, bool MASKED, bool SWZ = false>
__device__ __forceinline__ void p2_group(const u32 (&w)[K + 1][VEC], const u32 (&rad)[K + 1], u32 mul, u32 plow, u32 plow_mul,
                                         u32 *hist, int lim, u32 k30, u32 k28, u32 k26) {
    constexpr int C = K + 1;                  // columns, child last
    constexpr int C1 = C > 4 ? C - 4 : 0;     // columns of the high group
#pragma unroll
    for (int wd = 0; wd < VEC; ++wd) {
        u32 hi[4], lo[4];
#if BIC_P2_PAIRS
        // Two neighbouring columns of a group are combined while their values still sit in nibbles:
        // E = W & 0x33333333 holds rows 4j and 4j+2 of byte j in its two nibbles, O = (W >> 2) & ... rows
        // 4j+1 and 4j+3; E_a * rad_b + E_b stays below 16 per nibble (states <= 4), so one IMAD pairs
        // eight rows, and the pair is opened into byte lanes with 6 instead of 2 x 7 ALU-pipe
        // instructions (per 16 rows of a 6-column family: 68 instead of 74, IMADs 36 instead of 42).
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int a0 = g == 0 ? 0 : C1, a1 = g == 0 ? C1 : C;   // columns of the group (hi, then lo)
            u32 (&acc)[4] = g == 0 ? hi : lo;
#pragma unroll
            for (int a = a0; a < a1; a += 2) {
                u32 u[4];
                u32 radp;   // radix of what u holds
                if (a + 1 < a1) {
                    const u32 Wa = w[a][wd], Wb = w[a + 1][wd];
                    const u32 pe = (Wa & 0x33333333u) * rad[a + 1] + (Wb & 0x33333333u);
                    const u32 po = ((Wa >> 2) & 0x33333333u) * rad[a + 1] + ((Wb >> 2) & 0x33333333u);
                    u[0] = pe & 0x0f0f0f0fu;
                    u[1] = po & 0x0f0f0f0fu;
                    u[2] = (pe >> 4) & 0x0f0f0f0fu;
                    u[3] = (po >> 4) & 0x0f0f0f0fu;
                    radp = rad[a] * rad[a + 1];
                } else {
                    unpack2(w[a][wd], u, k30, k28, k26);
                    radp = rad[a];
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) acc[s] = (a == a0) ? u[s] : acc[s] * radp + u[s];
            }
        }
#else
#pragma unroll
        for (int a = 0; a < C; ++a) {
            u32 u[4];
            unpack2(w[a][wd], u, k30, k28, k26);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                if (a < C1) hi[s] = (a == 0) ? u[s] : hi[s] * rad[a] + u[s];
                else lo[s] = (a == C1) ? u[s] : lo[s] * rad[a] + u[s];
            }
        }
#endif
#if BIC_P2_DP4A
        // Byte lanes -> one counter offset per row with the integer dot product (IDP.4A, not an
        // ALU-pipe instruction): dp4a(lanes, weights) picks a row's byte(s) by the position of the
        // non-zero weights and scales them in the same instruction.  Without a high group the
        // weight is mul itself (<= 128): one IDP per row and nothing else.  With one, lo and hi are
        // interleaved (one PRMT per two rows), the weights are (1, plow) and the offset is
        // index * mul; plow = 256 (four low columns of four states) does not fit a byte weight, the
        // interleaved 16-bit lanes then are the index already.
        if (C1 == 0) {
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int B = 0; B < 4; ++B)
                    if (!MASKED || (16 * wd + 4 * B + s) < lim) {
                        const u32 o = __dp4a(lo[s], mul << (8 * B), 0u);
                        bump_off<false>(hist, SWZ ? swz_off(o) : o);
                    }
        } else if (plow < 256u) {
            const u32 wv = 1u | (plow << 8);
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const u32 v01 = __byte_perm(lo[s], hi[s], 0x5140u);   // (lo, hi) of rows B = 0, 1
                const u32 v23 = __byte_perm(lo[s], hi[s], 0x7362u);   // rows B = 2, 3
                const u32 off[4] = {__dp4a(v01, wv, 0u) * mul, __dp4a(v01, wv << 16, 0u) * mul,
                                    __dp4a(v23, wv, 0u) * mul, __dp4a(v23, wv << 16, 0u) * mul};
#pragma unroll
                for (int B = 0; B < 4; ++B)
                    if (!MASKED || (16 * wd + 4 * B + s) < lim) bump_off<false>(hist, SWZ ? swz_off(off[B]) : off[B]);
            }
        } else {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const u32 t01 = __byte_perm(lo[s], hi[s], 0x5140u) * mul;   // 16-bit lanes hi * 256 + lo
                const u32 t23 = __byte_perm(lo[s], hi[s], 0x7362u) * mul;
                const u32 off[4] = {t01 & 0xffffu, t01 >> 16, t23 & 0xffffu, t23 >> 16};
#pragma unroll
                for (int B = 0; B < 4; ++B)
                    if (!MASKED || (16 * wd + 4 * B + s) < lim) bump_off<false>(hist, SWZ ? swz_off(off[B]) : off[B]);
            }
        }
#else
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            u32 t01 = __byte_perm(lo[s], 0u, 0x4140u) * mul;   // rows B = 0, 1 in 16-bit lanes
            u32 t23 = __byte_perm(lo[s], 0u, 0x4342u) * mul;   // rows B = 2, 3
            if (C1 > 0) {
                t01 += __byte_perm(hi[s], 0u, 0x4140u) * plow_mul;
                t23 += __byte_perm(hi[s], 0u, 0x4342u) * plow_mul;
            }
            const u32 off[4] = {t01 & 0xffffu, t01 >> 16, t23 & 0xffffu, t23 >> 16};
#pragma unroll
            for (int B = 0; B < 4; ++B)
                if (!MASKED || (16 * wd + 4 * B + s) < lim) bump_off<false>(hist, SWZ ? swz_off(off[B]) : off[B]);
        }
#endif
    }
}

// [b0, b1): the CTA's slice in 512-row blocks (128 bytes of a packed column)
template <int K, int THREADS, int VEC, bool SWZ = false>
__device__ __forceinline__ void count_rows_p2(const FamMeta &m, const uint8_t *__restrict__ data2, long long stride2,
                                              long long N, long long b0, long long b1, u32 *hist, u32 k30, u32 k28, u32 k26) {
    constexpr int C = K + 1;
    constexpr int C1 = C > 4 ? C - 4 : 0;
    constexpr int ROWS = 16 * VEC;            // rows of one thread-iteration
    const uint8_t *cp[C];
    u32 rad[C];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data2 + (long long)m.par[a] * stride2;
        rad[a] = m.rad[a];
    }
    cp[K] = data2 + (long long)m.node * stride2;
    rad[K] = (u32)m.r;
    u32 plow = 1;
#pragma unroll
    for (int a = C1; a < C; ++a) plow *= rad[a];
    const u32 mul = m.mul, plow_mul = plow * mul;
    const long long g0 = b0 * (512 / ROWS), g1 = min(b1 * (512 / ROWS), (N + ROWS - 1) / ROWS);
    // (A variant in which the warps draw their row chunks from a shared-memory ticket instead of this
    // fixed stride, to even out the arrival at the barrier behind the loop, was 3 % slower — and the
    // loop shape that served both variants cost the fixed stride 7 %: 61.1 vs 56.8 ms per launch.)
    for (long long g = g0 + threadIdx.x; g < g1; g += THREADS) {
        u32 w[C][VEC];
#pragma unroll
        for (int a = 0; a < C; ++a) P2Load<VEC>::ld(cp[a] + g * (4 * VEC), w[a]);
        const long long row0 = g * ROWS;
        if (row0 + ROWS <= N) p2_group<K, VEC, false, SWZ>(w, rad, mul, plow, plow_mul, hist, ROWS, k30, k28, k26);
        else p2_group<K, VEC, true, SWZ>(w, rad, mul, plow, plow_mul, hist, (int)(N - row0), k30, k28, k26);
    }
}

// Packed path with two 64-row groups in flight per thread (families of <= 3 columns; BIC_P2_TWO=1,
// experiment).  Meant for packed datasets that do not fit L2 (pigs-shaped: 1.4 GB); measured 5 %
// slower there (class-0 launch 1.033 -> 1.087 ms, three runs each), so it is off by default.
template <int K, int THREADS>
__device__ __forceinline__ void count_rows_p2_two(const FamMeta &m, const uint8_t *__restrict__ data2, long long stride2,
                                                  long long N, long long b0, long long b1, u32 *hist, u32 k30, u32 k28, u32 k26) {
    constexpr int C = K + 1;
    constexpr int C1 = C > 4 ? C - 4 : 0;
    const uint8_t *cp[C];
    u32 rad[C];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data2 + (long long)m.par[a] * stride2;
        rad[a] = m.rad[a];
    }
    cp[K] = data2 + (long long)m.node * stride2;
    rad[K] = (u32)m.r;
    u32 plow = 1;
#pragma unroll
    for (int a = C1; a < C; ++a) plow *= rad[a];
    const u32 mul = m.mul, plow_mul = plow * mul;
    const long long g0 = b0 * 8, g1 = min(b1 * 8, (N + 63) / 64);
    for (long long g = g0 + threadIdx.x; g < g1; g += 2 * THREADS) {
        const long long gb = g + THREADS;
        const bool second = gb < g1;
        u32 w[C][4], wb[C][4];
#pragma unroll
        for (int a = 0; a < C; ++a) P2Load<4>::ld(cp[a] + g * 16, w[a]);
        if (second) {
#pragma unroll
            for (int a = 0; a < C; ++a) P2Load<4>::ld(cp[a] + gb * 16, wb[a]);
        }
        if (g * 64 + 64 <= N) p2_group<K, 4, false>(w, rad, mul, plow, plow_mul, hist, 64, k30, k28, k26);
        else p2_group<K, 4, true>(w, rad, mul, plow, plow_mul, hist, (int)(N - g * 64), k30, k28, k26);
        if (second) {
            if (gb * 64 + 64 <= N) p2_group<K, 4, false>(wb, rad, mul, plow, plow_mul, hist, 64, k30, k28, k26);
            else p2_group<K, 4, true>(wb, rad, mul, plow, plow_mul, hist, (int)(N - gb * 64), k30, k28, k26);
        }
    }
}

template <int THREADS, int VEC, bool SWZ = false>
__device__ __forceinline__ void count_rows_p2_k(const FamMeta &m, const uint8_t *__restrict__ data2, long long stride2,
                                                long long N, long long b0, long long b1, u32 *hist, u32 k30, u32 k28, u32 k26) {
    switch (m.k) {
        case 0: count_rows_p2<0, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
        case 1: count_rows_p2<1, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
        case 2: count_rows_p2<2, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
        case 3: count_rows_p2<3, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
        case 4: count_rows_p2<4, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
        case 5: count_rows_p2<5, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
        default: count_rows_p2<6, THREADS, VEC, SWZ>(m, data2, stride2, N, b0, b1, hist, k30, k28, k26); break;
    }
}

// Any number of parents (k > 6): columns are walked one at a time.
template <bool GLOBAL, int THREADS, bool RANGE = false>
__device__ __forceinline__ void count_rows_any(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                               long long N, long long v0, long long v1, u32 *hist) {
    const uint8_t *child = data + (long long)m.node * stride;
    for (long long v = v0 + threadIdx.x; v < v1; v += THREADS) {
        u32 off[16];
#pragma unroll
        for (int b = 0; b < 16; ++b) off[b] = 0;
        for (int a = 0; a <= m.k; ++a) {
            uint4 w = ld_stream_v4((a < m.k ? data + (long long)m.par[a] * stride : child) + v * 16);
            u32 rad = a < m.k ? m.rad[a] : (u32)m.r;
            const u32 ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int b = 0; b < 4; ++b) off[i * 4 + b] = off[i * 4 + b] * rad + ((ws[i] >> (8 * b)) & 0xffu);
        }
#pragma unroll
        for (int b = 0; b < 16; ++b) off[b] *= m.mul;
        bump16<GLOBAL, RANGE>(hist, off, v * 16, N, m.lo4, m.span4);
    }
}

// Sum the R lane replicas of every cell into hist[0 .. cells).  Thread t owns cells t,
// t+THREADS, ...; sums are parked in registers across the barrier because the compacted table
// overlaps the replicated one.  Replica order is rotated by the thread index so the R reads of
// a warp fall into different banks.
template <int THREADS>
__device__ __forceinline__ void compact_replicas(u32 *hist, u32 cells, u32 R) {
    u32 loc[REPL_MAX_PER_THREAD];
#pragma unroll
    for (int i = 0; i < REPL_MAX_PER_THREAD; ++i) {
        u32 c = threadIdx.x + i * THREADS;
        u32 s = 0;
        if (c < cells)
            for (u32 r = 0; r < R; ++r) s += hist[c * R + ((r + threadIdx.x) & (R - 1))];
        loc[i] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < REPL_MAX_PER_THREAD; ++i) {
        u32 c = threadIdx.x + i * THREADS;
        if (c < cells) hist[c] = loc[i];
    }
    __syncthreads();
}

// The fp64 reduce always runs on RED_LANES = 256 virtual lanes, whatever the block size: lane t
// owns parent configurations t, t+256, ... and the partial sums are combined in a fixed order
// (shuffle tree per warp, then the 8 warps in index order).  A family's log-likelihood is
// therefore bit-identical for every kernel class, slice count and for the row-sharded path.
constexpr int RED_LANES = 256;
constexpr u32 STAGE_WORDS = 6144;   // staging buffer of the kernels that reduce HBM tables only (24 KB)

// k3.  Virtual lane v (0..255) owns parent configurations v, v+256, ... and folds their terms
// into its accumulator in (j, x) order; a block of fewer than 256 threads runs two virtual lanes
// per thread (a warp always holds 32 consecutive virtual lanes), so the bits do not depend on
// the block size.  `row` is a functor (row pointer, accumulator) that adds one parent
// configuration's terms.
//
// Tables in shared memory are read in place.  Tables in HBM are staged through `stage` (cap
// words of shared memory) in pieces of cap / r consecutive configurations, loaded by the whole
// block with coalesced, independent loads: one CTA walking a 200 k-cell table with dependent
// 4-byte loads (the first version) spent ~150 us in L2 latency per family.  The order in which a
// lane meets its configurations is the same either way.
// HBM-table reduce options: `wb` (nullable) also receives a copy of the table as it streams
// through shared memory (row-sharded runs: the owner's summed table goes back into the arena when
// derived families need it as their donor).
struct TabSrc {
    u32 *wb;
};

template <bool FROM_GLOBAL, class RowFn>
__device__ __forceinline__ double reduce_rows(const u32 *tab, u32 q, int r, double *sh, u32 *stage, u32 cap, RowFn row,
                                              TabSrc ts = TabSrc{nullptr}) {
    double acc[2] = {0.0, 0.0};
    if (!FROM_GLOBAL) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const u32 vl = threadIdx.x + u * blockDim.x;
            if (vl < RED_LANES && (u == 0 || blockDim.x < RED_LANES))
                for (u32 j = vl; j < q; j += RED_LANES) row(tab + (size_t)j * r, acc[u]);
        }
    } else {
        const u32 J = cap / (u32)r;   // configurations per piece (r <= 255, cap >= 2048)
        for (u32 base = 0; base < q; base += J) {
            const u32 cnt = min(J, q - base), ncell = cnt * (u32)r;
            const u32 *src = tab + (size_t)base * r;
            __syncthreads();          // the previous piece has been consumed
            u32 i = threadIdx.x;
            u32 *wb = ts.wb ? ts.wb + (size_t)base * r : nullptr;
            for (; i + 7 * blockDim.x < ncell; i += 8 * blockDim.x) {
                u32 v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __ldcg(src + i + e * blockDim.x);
#pragma unroll
                for (int e = 0; e < 8; ++e) stage[i + e * blockDim.x] = v[e];
                if (wb) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) wb[i + e * blockDim.x] = v[e];
                }
            }
            for (; i < ncell; i += blockDim.x) {
                const u32 v = __ldcg(src + i);
                stage[i] = v;
                if (wb) wb[i] = v;
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const u32 vl = threadIdx.x + u * blockDim.x;
                if (vl < RED_LANES && (u == 0 || blockDim.x < RED_LANES))
                    for (u32 j = base + ((vl - base) & (RED_LANES - 1)); j < base + cnt; j += RED_LANES)
                        row(stage + (size_t)(j - base) * r, acc[u]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const u32 vl = threadIdx.x + u * blockDim.x;
        if (vl < RED_LANES && (u == 0 || blockDim.x < RED_LANES)) {
            double t = acc[u];
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
            if ((vl & 31) == 0) sh[vl >> 5] = t;
        }
    }
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < RED_LANES / 32; ++w) t += sh[w];
    return t;
}

// sum_{j,x: c>0} c * ln(c / N_ij)
template <bool FROM_GLOBAL>
__device__ __forceinline__ double family_loglik(const u32 *tab, u32 q, int r, double *sh, u32 *stage, u32 cap,
                                                TabSrc ts = TabSrc{nullptr}) {
    return reduce_rows<FROM_GLOBAL>(tab, q, r, sh, stage, cap, [r](const u32 *row, double &acc) {
        u32 nij = 0;
        for (int x = 0; x < r; ++x) nij += row[x];
        if (nij) {
            const double dn = (double)nij;
            for (int x = 0; x < r; ++x) {
                const u32 c = row[x];
                if (c) acc += (double)c * log((double)c / dn);
            }
        }
    }, ts);
}

// Bayesian-Dirichlet family term (bnlearn "bde" = BDeu, "k2") on the same counts, same lane order:
// sum_j [ lgamma(a_ij) - lgamma(a_ij + N_ij) + sum_x ( lgamma(a_ijk + c) - lgamma(a_ijk) ) ].
template <bool FROM_GLOBAL>
__device__ __forceinline__ double family_bd(const u32 *tab, u32 q, int r, double a_ij, double a_ijk, double *sh,
                                            u32 *stage, u32 cap, TabSrc ts = TabSrc{nullptr}) {
    const double lg_ij = lgamma(a_ij), lg_ijk = lgamma(a_ijk);
    return reduce_rows<FROM_GLOBAL>(tab, q, r, sh, stage, cap, [=](const u32 *row, double &acc) {
        u32 nij = 0;
        double s = 0.0;
        for (int x = 0; x < r; ++x) {
            const u32 c = row[x];
            nij += c;
            if (c) s += lgamma(a_ijk + (double)c) - lg_ijk;
        }
        if (nij) acc += (lg_ij - lgamma(a_ij + (double)nij)) + s;
    }, ts);
}

// The cached family term: log-likelihood (penalty applied at gather time) or a BD score.
// FROM_GLOBAL: `stage` / `cap` = shared-memory staging buffer (see reduce_rows).
template <bool FROM_GLOBAL>
__device__ __forceinline__ double family_term(const CountArgs &a, const u32 *tab, const FamMeta &m, double *sh,
                                              u32 *stage = nullptr, u32 cap = 0, TabSrc ts = TabSrc{nullptr}) {
    if (a.bd_mode == 0) return family_loglik<FROM_GLOBAL>(tab, m.q, m.r, sh, stage, cap, ts);
    double a_ijk = a.bd_mode == 1 ? a.iss / ((double)m.q * (double)m.r) : 1.0;
    return family_bd<FROM_GLOBAL>(tab, m.q, m.r, a_ijk * (double)m.r, a_ijk, sh, stage, cap, ts);
}

// A CTA's slice of the rows in 512-row blocks: 512 bytes of a uint8 column, 128 bytes of a 2-bit
// packed one, so every warp load covers whole cache lines (ncu showed 5.2 data-pipe wavefronts per
// 512-byte load with unaligned slices instead of 4).
__device__ __forceinline__ void slice_blocks(long long N, int slice, int S, long long &b0, long long &b1) {
    const long long nb = (N + 511) >> 9;
    b0 = nb * slice / S;
    b1 = (slice + 1 == S) ? nb : nb * (slice + 1) / S;
}

// Fused reduce-scatter step of a row-sharded run: the CTA that holds the family's complete LOCAL
// table (shared memory when the family has one slice, else the merged HBM table) stores it into
// slot `rank` of the owner rank's exchange buffer with coalesced peer stores over NVLink.
template <int THREADS, bool FROM_SHARED>
__device__ __forceinline__ void push_table(const CountArgs &a, int j, const u32 *src, u32 cells) {
    u32 *dst = a.peer[a.owner[j]] + (size_t)a.rank * a.xcap + a.xoff[j];   // 16-byte aligned (k_owner_offsets)
    if (FROM_SHARED) {   // the shared-memory table starts 16-byte aligned: 16 bytes per store, 512 per warp
        const u32 n4 = cells >> 2;
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (u32 c = threadIdx.x; c < n4; c += THREADS) d4[c] = s4[c];
        for (u32 c = (n4 << 2) + threadIdx.x; c < cells; c += THREADS) dst[c] = src[c];
    } else {
        for (u32 c = threadIdx.x; c < cells; c += THREADS) dst[c] = __ldcg(src + c);
    }
}

// Class 3 in sub-range passes, step 1: the cell index of every row is computed ONCE (k_cells) and
// parked in HBM scratch, 4 bytes per row; the P passes of k_count<.., false, true> then read it back
// (from L2: the passes of a slice are neighbours in the grid) and only compare and increment.  The
// round-1 passes recomputed the mixed-radix index of every row in each pass and threw (P - 1) / P of
// them away (ncu: issue-active 76 %, 0.15 of the HBM peak).
template <int K>
__device__ __forceinline__ void cells_rows_k(const FamMeta &m, const uint8_t *__restrict__ data, long long stride, long long v,
                                             u32 *out) {
    uint4 w[K + 1];
    u32 rad[K + 1];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        w[a] = ld_stream_v4(data + (long long)m.par[a] * stride + v * 16);
        rad[a] = m.rad[a];
    }
    w[K] = ld_stream_v4(data + (long long)m.node * stride + v * 16);
    rad[K] = (u32)m.r;
    u32 cell[16];
    cells_u32<K>(w, rad, 1u, cell);
    uint4 *o = reinterpret_cast<uint4 *>(out + v * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = make_uint4(cell[4 * i], cell[4 * i + 1], cell[4 * i + 2], cell[4 * i + 3]);
}

// grid = (CTAs per family, class-3 families); every thread turns 16 rows per iteration into 16 cells
__global__ void __launch_bounds__(256) k_cells(CountArgs a) {
    __shared__ FamMeta m;
    const int jj = blockIdx.y;
    const int j = a.jobs[jj];
    if (threadIdx.x == 0) decode_family(a.keys + (a.key_base + j) * (long long)(a.W64 + 1), a.W64, a.card, m);
    __syncthreads();
    u32 *out = a.cellbuf + (size_t)jj * (size_t)a.stride;
    const long long nvec = (a.N + 15) >> 4;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
        switch (m.k) {
            case 0: cells_rows_k<0>(m, a.data, a.stride, v, out); break;
            case 1: cells_rows_k<1>(m, a.data, a.stride, v, out); break;
            case 2: cells_rows_k<2>(m, a.data, a.stride, v, out); break;
            case 3: cells_rows_k<3>(m, a.data, a.stride, v, out); break;
            case 4: cells_rows_k<4>(m, a.data, a.stride, v, out); break;
            case 5: cells_rows_k<5>(m, a.data, a.stride, v, out); break;
            case 6: cells_rows_k<6>(m, a.data, a.stride, v, out); break;
            default: {   // k > 6: columns one at a time
                u32 cell[16];
#pragma unroll
                for (int b = 0; b < 16; ++b) cell[b] = 0;
                for (int x = 0; x <= m.k; ++x) {
                    const uint4 w = ld_stream_v4(a.data + (long long)(x < m.k ? m.par[x] : m.node) * a.stride + v * 16);
                    const u32 rad = x < m.k ? m.rad[x] : (u32)m.r;
                    const u32 ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int b = 0; b < 4; ++b) cell[i * 4 + b] = cell[i * 4 + b] * rad + ((ws[i] >> (8 * b)) & 0xffu);
                }
                uint4 *o = reinterpret_cast<uint4 *>(out + v * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = make_uint4(cell[4 * i], cell[4 * i + 1], cell[4 * i + 2], cell[4 * i + 3]);
            }
        }
    }
}

// step 2 of a pass: 16 parked cells per thread-iteration; the CTA owns the cells [lo, lo + span)
template <int THREADS>
__device__ __forceinline__ void count_rows_cells(const u32 *__restrict__ cells, long long N, long long v0, long long v1, u32 *hist,
                                                 u32 lo, u32 span) {
    for (long long v = v0 + threadIdx.x; v < v1; v += THREADS) {
        const uint4 *p = reinterpret_cast<const uint4 *>(cells + v * 16);
        uint4 c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) c[i] = __ldcg(p + i);
        const u32 cs[16] = {c[0].x, c[0].y, c[0].z, c[0].w, c[1].x, c[1].y, c[1].z, c[1].w,
                            c[2].x, c[2].y, c[2].z, c[2].w, c[3].x, c[3].y, c[3].z, c[3].w};
        const int nv = v * 16 + 16 <= N ? 16 : (int)(N - v * 16);
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            const u32 d = cs[b] - lo;
            if (b < nv && d < span) bump_off<false>(hist, d << 2);
        }
    }
}

// RANGE (class 3 when the rows dwarf the table): the table does not fit one CTA's shared memory,
// so a (family, slice) is counted in P passes; pass p keeps cells [p * span, (p + 1) * span) in
// shared memory, streams the slice and skips the rows whose cell lies elsewhere.  The passes of a
// slice are neighbours in the grid, so all but the first read the rows from L2.  That trades
// P x the streaming for shared-memory atomics instead of one L2 atomic per row (measured
// 0.09-0.19 T/s for the whole GPU).  k_count_cluster (one pass, table spread over a thread-block
// cluster) and parked cell indices (k_cells) were both measured slower and are off by default.
#ifndef BIC_C0_MINBLOCKS
#define BIC_C0_MINBLOCKS 4   // 256-thread CTAs per SM the register allocation is sized for (experiment: 3 -> 80 registers)
#endif
template <int THREADS, bool GLOBAL, bool RANGE = false>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? BIC_C0_MINBLOCKS : THREADS > 256 ? 1024 / THREADS : 1) k_count(CountArgs a) {   // 64 registers: 1024 threads per SM
    static_assert(!(GLOBAL && RANGE), "a sub-range table lives in shared memory");
    extern __shared__ __align__(128) u32 s_hist[];
    __shared__ FamMeta m;
    __shared__ double s_red[32];
    __shared__ int s_last;
    __shared__ __align__(8) u64 s_full[TMA_STAGES];
    __shared__ FamMetaC s_c;

    // slice-major item order: the CTAs resident at any moment work on the same row window of
    // the dataset, which the host sizes to stay L2-resident
    const bool listed = RANGE && a.items3 != nullptr;
    const int per_slice = RANGE ? (listed ? a.nitems3 : a.njobs * a.P) : a.njobs;
    const int slice = blockIdx.x / per_slice;
    const int in_slice = blockIdx.x - slice * per_slice;
    int pass = 0, j;
    if (listed) {
        const int2 it = a.items3[in_slice];
        j = it.x;
        pass = it.y;
    } else {
        pass = RANGE ? in_slice / a.njobs : 0;
        j = a.jobs[in_slice - pass * a.njobs];
    }
    if (a.meta) {   // one coalesced 128-byte read of the parked record
        if (threadIdx.x < 8) reinterpret_cast<uint4 *>(&s_c)[threadIdx.x] = __ldg(reinterpret_cast<const uint4 *>(a.meta + j) + threadIdx.x);
    } else if (threadIdx.x == 0) {
        decode_family(a.keys + (a.key_base + j) * (long long)(a.W64 + 1), a.W64, a.card, m);
    }
    __syncthreads();

    const u32 cells = a.meta ? s_c.cells : m.cells;
    if (a.tier_hi && (cells < a.tier_lo || cells > a.tier_hi)) return;   // the other launch over this class list counts the family
    // RANGE: this family's sub-ranges (generic cut, or runs of states of the first parent)
    RangePlan rp;
    rp.span = a.span; rp.passes = 1; rp.ns = 0;
    u32 low = 0;   // top split: cells below the first parent
    if (RANGE) {
        const int kk = a.meta ? s_c.k : m.k;
        const u32 rad0 = kk > 0 ? (a.meta ? (u32)s_c.rad[0] : m.rad[0]) : 1u;
        if (a.c3_u16 && kk <= 6) rp = range_plan(cells, kk, rad0, 2u * a.span, false);   // two 16-bit counters per word
        else rp = range_plan(cells, kk, rad0, a.span, a.topsplit != 0);
        low = rp.ns ? cells / rad0 : 0u;
    }
    const u32 lo = RANGE ? (u32)pass * rp.span : 0u;
    if (RANGE && lo >= cells) return;   // this family needs fewer passes than the largest of the launch
    const u32 span = RANGE ? min(rp.span, cells - lo) : cells;
    long long b0, b1;
    slice_blocks(a.N, slice, a.S, b0, b1);
    const long long nvec = (a.N + 15) >> 4;
    const long long v0 = b0 * 32, v1 = min(b1 * 32, nvec);   // 16-row vectors of the uint8 columns
    // Lane replicas (every thread computes the same R; thread 0 publishes it with the rest of the
    // decoded family before the barrier that follows the zeroing).  Replicas pay off only when the
    // row loop dwarfs zeroing + summing R tables: at least 16 rows per replicated counter.
    u32 R0 = 1;
    {
        const long long rows_here = (v1 - v0) * 16;
        // The packed path builds its counter offsets in 32 bits (p2_group, IDP.4A) unless the four low
        // columns have four states each; everything else carries them on 16-bit lanes: cells * R <= 16383.
        bool wide_off = false;
#if BIC_P2_DP4A
        if (!GLOBAL && !RANGE && a.data2 != nullptr) {
            const int kk = a.meta ? s_c.k : m.k;
            const bool sm = a.meta ? s_c.small != 0 : m.small != 0;
            if (sm && kk <= 6 && cells <= 16383u) {
                u32 plow = (u32)(a.meta ? s_c.r : m.r);
                for (int i = kk > 3 ? kk - 3 : 0; i < kk; ++i) plow *= a.meta ? (u32)s_c.rad[i] : m.rad[i];
                wide_off = kk < 4 || plow < 256u;
            }
        }
#endif
        if (!GLOBAL && !RANGE)
            while (R0 < 32 && cells * (R0 * 2) <= a.cap_words && (wide_off || cells * (R0 * 2) <= 16383u) &&   // 16-bit lane offsets
                   cells <= (u32)(REPL_MAX_PER_THREAD * THREADS) && (long long)cells * (R0 * 2) * 16 <= rows_here)
                R0 *= 2;
        // measured (ncu source counters, tools/microbench2): R = 32 -> 1.00 wavefront per warp
        // atomic, 16 -> 2.0, none -> ~2.6, but 8 -> 2.8 and 4 -> 3.3 because interleaving then
        // confines each lane to 4 or 8 banks.  So: 32, 16 or nothing.
        if (R0 < 16) R0 = 1;
    }
    if (a.meta && threadIdx.x < KMAX) {
        m.par[threadIdx.x] = s_c.par[threadIdx.x];
        m.rad[threadIdx.x] = s_c.rad[threadIdx.x];
    }
    if (threadIdx.x == 0) {
        if (a.meta) { m.k = s_c.k; m.node = s_c.node; m.r = s_c.r; m.q = s_c.q; m.cells = s_c.cells; m.small = s_c.small; }
        m.R = R0;
        m.mul = 4u * R0;
        m.lo4 = lo * 4u;
        m.span4 = span * 4u;
    }
    u32 *tab = (a.need && a.need[j]) ? a.arena + a.table_off[j] : nullptr;
    u32 *hist = GLOBAL ? tab : s_hist + (threadIdx.x & (R0 - 1));
    if (!GLOBAL)
        for (u32 c = threadIdx.x; c < ((RANGE && a.c3_u16) ? a.span : max(span * R0, (span + 31u) & ~31u)); c += THREADS)
            s_hist[c] = 0;   // whole 32-cell groups (swizzle); 16-bit mode: all a.span words
    __syncthreads();   // the decoded family, R and the zeroed table are visible
    const u32 R = m.R;

    // packed path: all columns <= 4 states, index * mul fits 16-bit lanes.  The choice must not
    // depend on the slice, hence no R here: R > 1 implies cells * R <= 16383 (replica selection above).
    const bool packed = !GLOBAL && !RANGE && a.data2 != nullptr && m.small && m.k <= 6 && cells <= 16383u;
    // un-replicated table on the packed path: spread the banks.  (On the uint8 path the same swizzle changed
    // nothing: diabetes-shaped classes 0 / 1 / 2 1.142 / 0.690 / 0.349 -> 1.145 / 0.682 / 0.358 ms; with 3 - 21
    // states per column the low index bits are spread already.)
    const bool swz = packed && R == 1 && a.swizzle && a.p2_vec == 4 && cells >= 64u;
    if (packed && swz) {
        count_rows_p2_k<THREADS, 4, true>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
    } else if (packed) {
        if (a.p2_two && m.k <= 2) {
            if (m.k == 0) count_rows_p2_two<0, THREADS>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
            else if (m.k == 1) count_rows_p2_two<1, THREADS>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
            else count_rows_p2_two<2, THREADS>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
        } else if (a.p2_vec == 4) count_rows_p2_k<THREADS, 4>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
        else if (a.p2_vec == 2) count_rows_p2_k<THREADS, 2>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
        else count_rows_p2_k<THREADS, 1>(m, a.data2, a.stride2, a.N, b0, b1, hist, a.k30, a.k28, a.k26);
    } else if (RANGE && a.cellbuf) {
        count_rows_cells<THREADS>(a.cellbuf + (size_t)(in_slice - pass * a.njobs) * (size_t)a.stride, a.N, v0, v1, hist, lo, span);
    } else if (RANGE && a.c3_u16 && m.k <= 6) {
        switch (m.k) {
            case 0: count_rows_r16<0, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
            case 1: count_rows_r16<1, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
            case 2: count_rows_r16<2, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
            case 3: count_rows_r16<3, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
            case 4: count_rows_r16<4, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
            case 5: count_rows_r16<5, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
            default: count_rows_r16<6, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, lo, span, a.span, tab); break;
        }
    } else if (RANGE && rp.ns) {
        const u32 s0 = (u32)pass * rp.ns, nsh = span / low;   // the last pass may hold fewer states
        switch (m.k) {
            case 1: count_rows_top<1, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, s0, nsh, low * 4u); break;
            case 2: count_rows_top<2, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, s0, nsh, low * 4u); break;
            case 3: count_rows_top<3, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, s0, nsh, low * 4u); break;
            case 4: count_rows_top<4, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, s0, nsh, low * 4u); break;
            case 5: count_rows_top<5, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, s0, nsh, low * 4u); break;
            default: count_rows_top<6, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, s0, nsh, low * 4u); break;
        }
    } else if (!GLOBAL && !RANGE && THREADS <= 512 && a.u8_narrow && m.k <= 6) {
        switch (m.k) {
            case 0: count_rows_narrow_mode<0, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
            case 1: count_rows_narrow_mode<1, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
            case 2: count_rows_narrow_mode<2, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
            case 3: count_rows_narrow_mode<3, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
            case 4: count_rows_narrow_mode<4, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
            case 5: count_rows_narrow_mode<5, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
            default: count_rows_narrow_mode<6, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist); break;
        }
    } else if (!GLOBAL && !RANGE && THREADS <= 512 && a.tma && m.k <= 6) {
        uint8_t *ring = reinterpret_cast<uint8_t *>(s_hist + a.cap_words);
        switch (m.k) {
            case 0: count_rows_tma_mode<0, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
            case 1: count_rows_tma_mode<1, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
            case 2: count_rows_tma_mode<2, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
            case 3: count_rows_tma_mode<3, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
            case 4: count_rows_tma_mode<4, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
            case 5: count_rows_tma_mode<5, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
            default: count_rows_tma_mode<6, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist, ring, s_full); break;
        }
    } else
    switch (m.k) {
        case 0: count_rows_mode<0, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist, a.u8_two); break;
        case 1: count_rows_mode<1, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist, a.u8_two); break;
        case 2: count_rows_mode<2, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist, a.u8_two); break;
        case 3: count_rows_mode<3, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist, a.u8_two); break;
        case 4: count_rows_mode<4, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist); break;
        case 5: count_rows_mode<5, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist); break;
        case 6: count_rows_mode<6, GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist); break;
        default: count_rows_any<GLOBAL, THREADS, RANGE>(m, a.data, a.stride, a.N, v0, v1, hist); break;
    }
    __syncthreads();
    if (!GLOBAL && !RANGE && R > 1) compact_replicas<THREADS>(s_hist, cells, R);
    if (!GLOBAL && !RANGE && swz) {   // back to cell order: a warp per 32-cell group, read all, then write
        const u32 lane = threadIdx.x & 31u;
        for (u32 g = threadIdx.x >> 5; g < (cells + 31u) >> 5; g += THREADS / 32) {
            const u32 v = s_hist[g * 32u + (lane ^ (g & 31u))];
            __syncwarp();
            s_hist[g * 32u + lane] = v;
        }
        __syncthreads();
    }

    const bool single = !GLOBAL && !RANGE && a.S == 1;   // the CTA's shared-memory table is the whole local table
    if (a.push && single) {   // row-sharded: straight from shared memory to the owner rank, no local HBM copy
        push_table<THREADS, true>(a, j, s_hist, cells);
        return;
    }
    if (RANGE && a.c3_u16 && m.k <= 6) {   // what the phases left in the 16-bit halves
        for (u32 c = threadIdx.x; c < a.span; c += THREADS) {
            const u32 v = s_hist[c], l16 = v & 0xffffu, h16 = v >> 16;
            if (l16) atomicAdd(tab + lo + c, l16);
            if (h16) atomicAdd(tab + lo + a.span + c, h16);
        }
    } else if (!GLOBAL && tab) {   // merge this slice's shared-memory table into the HBM table
        for (u32 c = threadIdx.x; c < span; c += THREADS) {
            u32 v = s_hist[c];
            if (v) atomicAdd(tab + lo + c, v);
        }
    }
    if (!a.reduce && !a.push) return;

    double ll;
    if (single) {
        ll = family_term<false>(a, s_hist, m, s_red);
    } else {
        const u32 parts = RANGE ? (u32)a.S * rp.passes : (u32)a.S;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(a.done + j, 1u) == parts - 1);
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (a.push) {   // the local table is complete: hand it to the owner rank
            push_table<THREADS, false>(a, j, tab, cells);
            return;
        }
        ll = family_term<true>(a, tab, m, s_red, s_hist, a.stage_words);
    }
    if (threadIdx.x == 0) {
        a.ll_out[a.out_base + j] = ll;
        a.np_out[a.out_base + j] = a.bd_mode ? 0.0 : (double)(m.r - 1) * (double)m.q;
    }
}

// ---------------------------------------------------------------------------------------
// Class 3 in ONE pass over the rows: a thread-block cluster of CL CTAs (2, 4 or 8; one CTA per SM)
// holds the table of a (family, slice) in distributed shared memory, cell c in CTA c % CL at word
// c / CL (interleaved, so that skewed distributions load the CTAs evenly).  Every CTA streams 1/CL
// of the slice, computes each row's cell once and increments the owning CTA's counter with
// red.shared::cluster over the SM-to-SM network.  The sub-range passes above recompute the index
// of every row in each of P passes and discard (P - 1) / P of them (ncu: issue-active 76 %, ALU
// 66 %, 0.15 of the HBM peak).
__device__ __forceinline__ u32 cluster_ctarank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one increment of the counter at shared-memory byte address `local` of CTA `rank` of this cluster
__device__ __forceinline__ void cluster_inc(u32 local, u32 rank) {
    u32 remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(rank));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(1u) : "memory");
}

template <int K, int THREADS>
__device__ __forceinline__ void count_rows_cluster(const FamMeta &m, const uint8_t *__restrict__ data, long long stride,
                                                   long long N, long long v0, long long v1, u32 hist_addr, u32 clmask,
                                                   u32 clshift) {
    const uint8_t *cp[K + 1];
    u32 rad[K + 1];
#pragma unroll
    for (int a = 0; a < K; ++a) {
        cp[a] = data + (long long)m.par[a] * stride;
        rad[a] = m.rad[a];
    }
    cp[K] = data + (long long)m.node * stride;
    rad[K] = (u32)m.r;
    for (long long v = v0 + threadIdx.x; v < v1; v += THREADS) {
        uint4 w[K + 1];
#pragma unroll
        for (int a = 0; a <= K; ++a) w[a] = ld_stream_v4(cp[a] + v * 16);
        u32 cell[16];
        cells_u32<K>(w, rad, 1u, cell);
        const int nv = v * 16 + 16 <= N ? 16 : (int)(N - v * 16);
#pragma unroll
        for (int b = 0; b < 16; ++b)
            if (b < nv) cluster_inc(hist_addr + ((cell[b] >> clshift) << 2), cell[b] & clmask);
    }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_count_cluster(CountArgs a, int CL) {
    extern __shared__ u32 s_hist[];
    __shared__ FamMeta m;
    __shared__ double s_red[32];
    __shared__ int s_last;
    const u32 rank = cluster_ctarank();
    const int item = blockIdx.x / CL;            // (slice, family), slice-major
    const int slice = item / a.njobs;
    const int j = a.jobs[item - slice * a.njobs];
    if (threadIdx.x == 0) decode_family(a.keys + (a.key_base + j) * (long long)(a.W64 + 1), a.W64, a.card, m);
    __syncthreads();
    const u32 cells = m.cells;
    const u32 clshift = 31 - __clz(CL), clmask = (u32)CL - 1u;
    const u32 words = (cells + clmask - rank) >> clshift;   // cells c < cells with c % CL == rank
    for (u32 c = threadIdx.x; c < words; c += THREADS) s_hist[c] = 0;
    cluster_sync_all();                                     // every CTA's table is zeroed before the first remote increment

    long long b0, b1;
    slice_blocks(a.N, slice, a.S, b0, b1);
    const long long bq0 = b0 + (b1 - b0) * rank / CL, bq1 = b0 + (b1 - b0) * (rank + 1) / CL;   // this CTA's share of the slice
    const long long nvec = (a.N + 15) >> 4;
    const long long v0 = bq0 * 32, v1 = min(bq1 * 32, nvec);
    const u32 hist_addr = (u32)__cvta_generic_to_shared(s_hist);
    switch (m.k) {
        case 0: count_rows_cluster<0, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        case 1: count_rows_cluster<1, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        case 2: count_rows_cluster<2, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        case 3: count_rows_cluster<3, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        case 4: count_rows_cluster<4, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        case 5: count_rows_cluster<5, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        case 6: count_rows_cluster<6, THREADS>(m, a.data, a.stride, a.N, v0, v1, hist_addr, clmask, clshift); break;
        default: {   // k > 6: columns one at a time
            const uint8_t *child = a.data + (long long)m.node * a.stride;
            for (long long v = v0 + threadIdx.x; v < v1; v += THREADS) {
                u32 cell[16];
#pragma unroll
                for (int b = 0; b < 16; ++b) cell[b] = 0;
                for (int x = 0; x <= m.k; ++x) {
                    uint4 w = ld_stream_v4((x < m.k ? a.data + (long long)m.par[x] * a.stride : child) + v * 16);
                    const u32 rad = x < m.k ? m.rad[x] : (u32)m.r;
                    const u32 ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int b = 0; b < 4; ++b) cell[i * 4 + b] = cell[i * 4 + b] * rad + ((ws[i] >> (8 * b)) & 0xffu);
                }
                const int nv = v * 16 + 16 <= a.N ? 16 : (int)(a.N - v * 16);
#pragma unroll
                for (int b = 0; b < 16; ++b)
                    if (b < nv) cluster_inc(hist_addr + ((cell[b] >> clshift) << 2), cell[b] & clmask);
            }
        }
    }
    cluster_sync_all();   // all increments of the cluster have landed; nobody touches remote shared memory after this

    // this CTA's interleaved share of the table goes to the HBM table (always: the fp64 reduce needs it in one place)
    u32 *tab = a.arena + a.table_off[j];
    for (u32 c = threadIdx.x; c < words; c += THREADS) {
        const u32 v = s_hist[c];
        if (v) {
            if (a.S == 1) tab[(c << clshift) + rank] = v;   // the arena is zeroed; one writer per cell
            else atomicAdd(tab + (c << clshift) + rank, v);
        }
    }
    if (!a.reduce && !a.push) return;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.done + j, 1u) == (u32)(a.S * CL) - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (a.push) {
        push_table<THREADS, false>(a, j, tab, cells);
        return;
    }
    const double ll = family_term<true>(a, tab, m, s_red, s_hist, a.stage_words);
    if (threadIdx.x == 0) {
        a.ll_out[a.out_base + j] = ll;
        a.np_out[a.out_base + j] = a.bd_mode ? 0.0 : (double)(m.r - 1) * (double)m.q;
    }
}

// Owner side of the fused reduce-scatter: slot `rank` of the own exchange buffer becomes the sum of
// all `world` slots over the cells this rank owns (the tables lie back to back, so this is one
// element-wise pass, 16 bytes per thread-iteration).  Runs after the barrier that orders every
// rank's peer stores.
__global__ void __launch_bounds__(256) k_sum_slots(u32 *xchg, u64 xcap, int world, int rank, const Header *hdr) {
    const u64 nvec = (hdr->owned_max + 3) >> 2;   // upper bound of this rank's owned cells (tables are 4-cell aligned)
    uint4 *dst = reinterpret_cast<uint4 *>(xchg + (size_t)rank * xcap);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (u64)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int s = 0; s < world; ++s) {
            const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(xchg + (size_t)s * xcap) + i);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        dst[i] = acc;
    }
}

// k3 alone: reduce HBM tables of a row-sharded run.  Without the fused reduce-scatter (world 1, the
// caller wants every table, or the exchange buffer is too small) every rank holds the all-reduced
// tables in its arena and reduces all of them.  With it (a.push) a rank reduces only the families it
// owns, from the sum of the `world` partial tables the ranks pushed into its exchange buffer.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_reduce_tables(CountArgs a, int njobs) {
    __shared__ FamMeta m;
    __shared__ double s_red[32];
    __shared__ u32 s_stage[STAGE_WORDS];
    const int j = blockIdx.x;
    if (j >= njobs) return;
    if (a.donor && a.donor[j] >= 0) return;   // derived families are reduced by k_derive
    if (a.push && a.owner[j] != a.rank) return;
    if (threadIdx.x == 0) decode_family(a.keys + (a.key_base + j) * (long long)(a.W64 + 1), a.W64, a.card, m);
    __syncthreads();
    double ll;
    if (a.push) {   // slot `rank` of the own exchange buffer holds the sum over all ranks (k_sum_slots)
        TabSrc ts{a.writeback ? a.arena + a.table_off[j] : nullptr};
        ll = family_term<true>(a, a.peer[a.rank] + (size_t)a.rank * a.xcap + a.xoff[j], m, s_red, s_stage, STAGE_WORDS, ts);
    } else {
        ll = family_term<true>(a, a.arena + a.table_off[j], m, s_red, s_stage, STAGE_WORDS);
    }
    if (threadIdx.x == 0) {
        a.ll_out[a.out_base + j] = ll;
        a.np_out[a.out_base + j] = a.bd_mode ? 0.0 : (double)(m.r - 1) * (double)m.q;
    }
}

// Derived families.  The donor's table is the joint table over a superset of {y} + Q in the
// donor's own axis order (its parents ascending, its child last).  Target cell t = (digits over
// Q ascending, then y) maps to the donor offset sum_v digit_v * stride_donor(v); the donor's
// extra axes are summed out.  One launch per level (number of parents), high to low, so a donor
// that is itself derived is complete (level_list = the families of this level).  A family is shared out over up to DERIVE_CHUNKS CTAs (one per
// ~4096 donor cells read: a 9261-cell table summed out of a 194 k-cell donor by a single CTA
// took 300 us of dependent L2 loads); the last chunk to finish (atomic ticket) runs the usual
// fp64 reduce.
constexpr int DERIVE_CHUNKS = 16;
template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_derive(CountArgs a, const int *__restrict__ level_list, int dch,
                                                    const u32 *__restrict__ cells_arr) {
    __shared__ FamMeta m, md;
    __shared__ double s_red[32];
    __shared__ u32 s_stage[STAGE_WORDS];
    __shared__ u32 s_stride[KMAX + 1];   // donor stride of own axis q (parents 0..k-1, child k)
    __shared__ u32 s_xstride[KMAX + 1], s_xrad[KMAX + 1];   // the donor's extra axes
    __shared__ u32 s_nx, s_xcells;
    __shared__ int s_last;
    const int j = level_list[blockIdx.x / dch];
    if (a.push && a.owner[j] != a.rank) return;   // the donor's global table lives on the owner rank only
    const u32 chunk = blockIdx.x % dch;
    const u32 nchunks = min((u32)dch, max(1u, cells_arr[a.donor[j]] / 4096u));
    if (chunk >= nchunks) return;
    if (threadIdx.x == 0) {
        const long long Wk = a.W64 + 1;
        const u64 *key = a.keys + (a.key_base + j) * Wk;
        {
            decode_family(key, a.W64, a.card, m);
            decode_family(a.keys + (a.key_base + a.donor[j]) * Wk, a.W64, a.card, md);
            for (int q = 0; q <= m.k; ++q) s_stride[q] = 0;
            u32 nx = 0, xcells = 1;
            // walk the donor's axes from the fastest (its child) to the slowest (first parent)
            u32 st = 1;
            for (int p = md.k; p >= 0; --p) {
                int v = p == md.k ? md.node : md.par[p];
                u32 rad = p == md.k ? (u32)md.r : md.rad[p];
                bool own = false;
                for (int q = 0; q <= m.k; ++q)
                    if ((q < m.k ? m.par[q] : m.node) == v) { s_stride[q] = st; own = true; }
                if (!own && rad > 1) {
                    s_xstride[nx] = st;
                    s_xrad[nx] = rad;
                    xcells *= rad;
                    ++nx;
                }
                st *= rad;
            }
            s_nx = nx;
            s_xcells = xcells;
        }
    }
    __syncthreads();
    const u32 *dt = a.arena + a.table_off[a.donor[j]];
    u32 *mt = a.arena + a.table_off[j];
    const u32 nx = s_nx, xcells = s_xcells, cells = m.cells;
    const u32 t0 = (u32)((u64)cells * chunk / nchunks), t1 = (u32)((u64)cells * (chunk + 1) / nchunks);
    const int k = m.k;
    // Few target cells, each the sum of many donor cells (a marginal of one or two variables out of
    // a 194 k-cell donor: 21 threads walked 9261 cells each, 136 us for a launch that moves 1 MB):
    // a warp per target cell, the lanes share out the extra cells.  Integer sums: same bits.
    const bool wide_sum = xcells >= 64u && (t1 - t0) * 8u <= (u32)THREADS;
    if (wide_sum) {
        const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        for (u32 t = t0 + warp; t < t1; t += THREADS / 32) {
            u32 rem = t / (u32)m.r;
            size_t src = (size_t)(t - rem * (u32)m.r) * s_stride[k];
            for (int q = k - 1; q >= 0; --q) {
                u32 nxt = rem / m.rad[q];
                src += (size_t)(rem - nxt * m.rad[q]) * s_stride[q];
                rem = nxt;
            }
            u32 s = 0;
            for (u32 e = lane; e < xcells; e += 32u) {
                u32 er = e;
                size_t o = src;
                for (u32 ax = 0; ax < nx; ++ax) {
                    u32 en = er / s_xrad[ax];
                    o += (size_t)(er - en * s_xrad[ax]) * s_xstride[ax];
                    er = en;
                }
                s += __ldcg(dt + o);
            }
            for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) mt[t] = s;
        }
    } else
    for (u32 t = t0 + threadIdx.x; t < t1; t += THREADS) {
        u32 rem = t / (u32)m.r;
        size_t src = (size_t)(t - rem * (u32)m.r) * s_stride[k];
        for (int q = k - 1; q >= 0; --q) {
            u32 nxt = rem / m.rad[q];
            src += (size_t)(rem - nxt * m.rad[q]) * s_stride[q];
            rem = nxt;
        }
        u32 s = 0;
        if (nx == 1) {   // one extra axis (the common case): independent strided loads
            const u32 *p = dt + src;
            const size_t xs = s_xstride[0];
            u32 e = 0;
            for (; e + 4 <= xcells; e += 4) {
                u32 v0 = __ldcg(p + (e + 0) * xs), v1 = __ldcg(p + (e + 1) * xs), v2 = __ldcg(p + (e + 2) * xs),
                    v3 = __ldcg(p + (e + 3) * xs);
                s += v0 + v1 + v2 + v3;
            }
            for (; e < xcells; ++e) s += __ldcg(p + e * xs);
        } else {
            for (u32 e = 0; e < xcells; ++e) {
                u32 er = e;
                size_t o = src;
                for (u32 ax = 0; ax < nx; ++ax) {
                    u32 en = er / s_xrad[ax];
                    o += (size_t)(er - en * s_xrad[ax]) * s_xstride[ax];
                    er = en;
                }
                s += __ldcg(dt + o);
            }
        }
        mt[t] = s;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(a.done + j, 1u) == nchunks - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double ll = family_term<true>(a, mt, m, s_red, s_stage, STAGE_WORDS);
    if (threadIdx.x == 0) {
        a.ll_out[a.out_base + j] = ll;
        a.np_out[a.out_base + j] = a.bd_mode ? 0.0 : (double)(m.r - 1) * (double)m.q;
    }
}

// Copy one family's table from the arena to the caller's layout (bic_count_families).
__global__ void k_copy_tables(const u32 *__restrict__ arena, const u64 *__restrict__ table_off,
                              const u32 *__restrict__ cells_arr, const long long *__restrict__ counts_off,
                              int *counts_out, Header *hdr) {
    int f = blockIdx.x;
    u32 cells = cells_arr[f];
    if ((long long)cells != counts_off[f + 1] - counts_off[f]) {
        if (threadIdx.x == 0) atomicOr(&hdr->err, 4u);
        return;
    }
    const u32 *src = arena + table_off[f];
    int *dst = counts_out + counts_off[f];
    for (u32 c = threadIdx.x; c < cells; c += blockDim.x) dst[c] = (int)src[c];
}

}  // namespace bic
