// Shared definitions for libbicgpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bic {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr int KMAX = 32;                   // parents with cardinality > 1 in one family (2^28 cells cap => <= 27)
constexpr int NMAX = 1024;                 // variables per dataset
constexpr int W64MAX = NMAX / 64;          // bitmask words per parent set
constexpr u32 ENT_EMPTY = 0u;              // hash-table entry: empty
constexpr u32 ENT_PENDING = 0x80000000u;   // entry = PENDING | instance index while a batch is being deduplicated
constexpr u64 MAX_CELLS = 1ull << 28;      // q*r limit of one count table (1 GiB of int32)
constexpr int DERIVE_LEVELS = 32;          // families with fewer parents than this may be derived from a superset

// Count-kernel classes by table size (cells = q*r).  Classes 0..2 keep the table in shared
// memory (one privatised histogram per CTA), class 3 counts straight into HBM with L2 atomics.
constexpr int NCLASS = 4;
constexpr u32 CLASS0_CELLS = 2048;         //   8 KB of int32 -> many CTAs per SM
constexpr u32 CLASS0_WORDS = 6144;         //  24 KB per class-0 CTA: room for lane replicas (count_kernels.cuh).  Swept 16..48 KB on
                                           //  B200: 24 KB is best; at 48 KB (192 KB per SM) the kernel is 15 % slower because too
                                           //  little L1 is left to land the in-flight streaming loads
constexpr u32 CLASS0_WORDS_PACKED = 12288; //  48 KB when all families stream the 2-bit packed copy (a quarter of the bytes in flight per
                                           //  row, so the L1 that is left suffices): 32 lane replicas up to 384 cells, 16 up to 768.
                                           //  Swept 24..64 KB on the alarm-shaped step: 61.1 / 60.9 / 58.2 / 59.3 ms at 24 / 32 / 48 / 64 KB
constexpr u32 CLASS1_CELLS = 12288;        //  48 KB
constexpr u32 CLASS2_CELLS = 49152;        // 192 KB (one CTA per SM)

// Device-written batch header, mirrored in pinned host memory once per sub-batch.
struct Header {
    u32 f_new;                 // families first seen in this sub-batch (exclusive-scan total)
    u32 class_count[NCLASS];   // of those, per count-kernel class
    u32 err;                   // bit 0: q*r over MAX_CELLS; bit 1: bad parent index
    u32 n_invalid;             // DAGs rejected (cyclic, self loop, labels not a permutation)
    u32 max_cells;             // largest count table among the new families
    u64 alg_bytes[NCLASS];     // sum (k+1)*N + 4*q*r over the new families, per class
    u64 class_cells[NCLASS];   // sum q*r over the new families, per class (what a slice merges into HBM)
    u64 table_cells;           // cells of all count tables that must live in HBM (scan total)
    u64 cells_all;             // cells of every described family: upper bound of table_cells, exact when all tables live in HBM
    u32 n_derived;             // new families whose table is marginalised from a counted superset
    u32 lvl_count[DERIVE_LEVELS];   // of those, by number of parents
    u32 lvl_cursor[DERIVE_LEVELS];  // device scratch: fill positions while the derived families are grouped by level
    u64 owned_max;             // row-sharded runs: most cells any one rank owns (must fit a slot of the exchange buffer)
    u32 max_passes3;           // most sub-range passes any class-3 family of the sub-batch needs (range_plan)
    u32 max_passes3_u16;       // the same with 16-bit counters (sub-ranges of twice the cells; k <= 6 only)
    u32 sum_passes3, sum_passes3_u16;   // sub-range passes of all class-3 families together: work items per row slice (k_range_items)
};

// How a class-3 family (table above one CTA's shared memory) is cut into sub-ranges that a CTA counts
// in shared memory, one pass over the row slice per sub-range.  Generic: sub-ranges of `span_max`
// cells, every pass computes the full mixed-radix cell of every row and keeps those inside.  Top
// split: when the table below the first (most significant) parent has at most 16383 cells, a
// sub-range is a run of `ns` states of that parent: the pass tests the parent's byte alone and
// builds the rest of the index on 16-bit packed lanes (two rows per IMAD), about half the
// instructions per row (about 11 against 15).  Taken only when it needs no more passes than the
// generic cut: with five passes against four (a 21^4-cell table) the diabetes-shaped class-3 launch
// measured 0.403 against 0.407 ms, the extra sweep over the rows eats the saving.
// describe_family (pass count of the launch) and k_count (the CTA's sub-range) both call this.
struct RangePlan {
    u32 span;     // cells per sub-range
    u32 passes;
    u32 ns;       // top split: states of the first parent per pass (0: generic)
};
__host__ __device__ inline RangePlan range_plan(u32 cells, int k, u32 rad0, u32 span_max, bool allow_top) {
    RangePlan p;
    p.span = span_max;
    p.passes = (cells + span_max - 1) / span_max;
    p.ns = 0;
    if (allow_top && k >= 1 && k <= 6 && rad0 > 1) {
        const u32 low = cells / rad0;
        if (low <= 16383u && low > 0) {
            u32 ns = span_max / low;
            const u32 np = (rad0 + ns - 1) / ns;
            ns = (rad0 + np - 1) / np;   // even out the passes
            if (np <= p.passes) {
                p.span = ns * low;
                p.passes = np;
                p.ns = ns;
            }
        }
    }
    return p;
}

__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// Family key: word 0 = node, words 1..W64 = parent bitmask.
__device__ __forceinline__ u64 hash_key(const u64 *key, int Wk) {
    u64 h = 0x9E3779B97F4A7C15ULL;
    for (int w = 0; w < Wk; ++w) h = mix64(h ^ key[w]) + 0x632BE59BD9B4E019ULL * (u64)(w + 1);
    return h;
}

__device__ __forceinline__ bool keys_equal(const u64 *a, const u64 *b, int Wk) {
    bool eq = true;
    for (int w = 0; w < Wk; ++w) eq = eq && (a[w] == b[w]);
    return eq;
}

// 128-bit streaming load of 16 consecutive rows of one state column (read once: keep out of L1).
__device__ __forceinline__ uint4 ld_stream_v4(const uint8_t *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

}  // namespace bic
