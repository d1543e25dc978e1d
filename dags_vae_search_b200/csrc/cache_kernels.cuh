// k5 (candidate -> family keys), acyclicity check, k4 (family-score cache: dedup / lookup /
// insert / gather-sum) and the small device scans that make family ids deterministic.
//
// Reference behaviour replaced: bnlearn.py:38-44 (relabel + adjacency serialisation),
// bnlearn_score.R:7-13,35 (parse adjacency, amat<- rejects cycles).  The cache itself has no
// reference equivalent (the reference recounts every family of every DAG).
#pragma once
#include "common.cuh"

namespace bic {

// ------------------------------------------------------------------------------- keys
// adj [B][n][n] uint8, row = parent, col = child.  One thread per (b, child i); consecutive
// threads read consecutive bytes of the same adjacency row.
__global__ void k_keys_adj(const uint8_t *__restrict__ adj, long long B, int n, int W64, u64 *keybuf,
                           uint8_t *dag_bad) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B * n) return;
    long long b = t / n;
    int i = (int)(t - b * n);
    const uint8_t *a = adj + b * (long long)n * n;
    u64 *key = keybuf + t * (W64 + 1);
    key[0] = (u64)i;
    for (int w = 0; w < W64; ++w) {
        u64 m = 0;
        int p1 = min(n, w * 64 + 64);
        for (int p = w * 64; p < p1; ++p)
            if (a[(long long)p * n + i]) m |= 1ull << (p & 63);
        key[1 + w] = m;
    }
    if (a[(long long)i * n + i]) dag_bad[b] = 1;  // self loop
}

// Parent lists in CSR.  node == nullptr: instance t is family (t / n, t % n) of a DAG batch.
__global__ void k_keys_csr(const long long *__restrict__ off, const int *__restrict__ parents,
                           const int *__restrict__ node, long long T, int n, int W64, u64 *keybuf,
                           uint8_t *dag_bad, Header *hdr) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int i = node ? node[t] : (int)(t % n);
    u64 *key = keybuf + t * (W64 + 1);
    for (int w = 0; w < W64; ++w) key[1 + w] = 0;
    bool bad = (i < 0 || i >= n);
    key[0] = bad ? 0 : (u64)i;
    long long e0 = off[t], e1 = off[t + 1];
    for (long long e = e0; e < e1; ++e) {
        int p = parents[e];
        if (p < 0 || p >= n || p == i) { bad = true; continue; }
        key[1 + (p >> 6)] |= 1ull << (p & 63);
    }
    if (bad) {
        if (node) atomicOr(&hdr->err, 2u);   // family API: hard error
        else dag_bad[t / n] = 1;             // DAG API: reject this DAG
    }
}

// Reference wire format (src/toolkit/labeled.py:132-154): vertex v carries label l_v (= BN
// variable), bit u of ebits[v] <=> edge vertex u -> vertex v, u < v.  bnlearn.py:38-42 relabels
// vertices to variables; here one thread per DAG does that and writes the n keys in variable
// order.  n <= 32, so one mask word.  LT: uint8 (legacy entry point) or uint16 (the reference's
// own label type, labeled.py:118).
template <typename LT>
__global__ void k_keys_wire(const LT *__restrict__ labels, const u32 *__restrict__ ebits,
                            long long B, int n, u64 *keybuf, uint8_t *dag_bad) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const LT *lab = labels + b * n;
    const u32 *eb = ebits + b * n;
    u32 seen = 0;
    for (int v = 0; v < n; ++v) {
        int l = lab[v];
        if (l < n) seen |= 1u << l;
    }
    u32 full = (n == 32) ? 0xffffffffu : ((1u << n) - 1u);
    if (seen != full) {  // bnlearn.py:35 asserts the labels are exactly 0..n-1
        dag_bad[b] = 1;
        for (int v = 0; v < n; ++v) { keybuf[(b * n + v) * 2] = (u64)v; keybuf[(b * n + v) * 2 + 1] = 0; }
        return;
    }
    for (int v = 0; v < n; ++v) {
        u32 e = eb[v] & ((v == 0) ? 0u : ((v >= 32) ? 0xffffffffu : ((1u << v) - 1u)));
        u64 m = 0;
        while (e) {
            int u = __ffs(e) - 1;
            e &= e - 1;
            m |= 1ull << lab[u];
        }
        long long t = b * n + lab[v];
        keybuf[t * 2] = (u64)lab[v];
        keybuf[t * 2 + 1] = m;
    }
}

// The same for any n <= 1024 (alarm-shaped n = 37, pigs-shaped n = 441): uint16 labels, EW =
// ceil(n / 32) edge words per vertex (bit u % 32 of word u / 32 of ebits[b][v] <=> edge u -> v).
// One block per DAG: the labels and a seen-bitmap sit in shared memory, thread v builds the
// parent mask of variable lab[v].
constexpr int WIRE_WIDE_THREADS = 128;
__global__ void __launch_bounds__(WIRE_WIDE_THREADS)
k_keys_wire_wide(const uint16_t *__restrict__ labels, const u32 *__restrict__ ebits, long long B, int n, int EW,
                 int W64, u64 *keybuf, uint8_t *dag_bad) {
    __shared__ uint16_t s_lab[NMAX];
    __shared__ u32 s_seen[NMAX / 32];
    __shared__ int s_ok;
    const long long b = blockIdx.x;
    const int Wk = W64 + 1;
    for (int w = threadIdx.x; w < NMAX / 32; w += blockDim.x) s_seen[w] = 0;
    if (threadIdx.x == 0) s_ok = 1;
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const int l = labels[b * n + v];
        s_lab[v] = (uint16_t)l;
        if (l < n) {
            if (atomicOr(&s_seen[l >> 5], 1u << (l & 31)) & (1u << (l & 31))) s_ok = 0;   // label used twice
        } else {
            s_ok = 0;
        }
    }
    __syncthreads();
    u64 *keys = keybuf + b * (long long)n * Wk;
    if (!s_ok) {   // bnlearn.py:35 asserts the labels are exactly 0..n-1
        if (threadIdx.x == 0) dag_bad[b] = 1;
        for (int v = threadIdx.x; v < n; v += blockDim.x) {
            keys[(long long)v * Wk] = (u64)v;
            for (int w = 0; w < W64; ++w) keys[(long long)v * Wk + 1 + w] = 0;
        }
        return;
    }
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        u64 m[W64MAX];
        for (int w = 0; w < W64; ++w) m[w] = 0;
        const u32 *eb = ebits + (b * n + v) * (long long)EW;
        for (int ew = 0; ew * 32 < v; ++ew) {
            u32 e = eb[ew];
            const int left = v - ew * 32;               // only u < v count (labeled.py:143-145)
            if (left < 32) e &= (1u << left) - 1u;
            while (e) {
                const int u = ew * 32 + __ffs(e) - 1;
                e &= e - 1;
                const int l = s_lab[u];
                m[l >> 6] |= 1ull << (l & 63);
            }
        }
        u64 *key = keys + (long long)s_lab[v] * Wk;
        key[0] = (u64)s_lab[v];
        for (int w = 0; w < W64; ++w) key[1 + w] = m[w];
    }
}

// ---------------------------------------------------------------------------- acyclic
// One warp per DAG: peel parentless vertices (Kahn) on bitmasks held in shared memory.
// dag_bad[b] in: self loop / bad index; out: also set when a cycle remains.
// One warp per DAG, or (WIDE, networks of more than 128 variables) one 256-thread block per DAG so
// that a peeling round tests at most four vertices per thread: 413 vertices on one warp cost
// 150 us per launch, latency-bound.
constexpr int ACYC_WARPS = 4;
constexpr int ACYC_WIDE_THREADS = 256;
constexpr u32 ACYC_SMEM_MAX = 48u << 10;   // WIDE: the parent masks of a DAG are staged in shared memory when they fit

// n <= 128: a warp per DAG, everything in registers.  Lane l holds the parent masks of vertices
// l, l + 32, l + 64, l + 96 and a copy of the alive set; a peeling round is a few logic
// instructions and one warp reduction per 32 vertices (the first version re-read the masks from
// global memory every round: 88 us for 100 k eleven-vertex DAGs, latency-bound).
__global__ void __launch_bounds__(ACYC_WARPS * 32)
k_acyclic_warp(const u64 *__restrict__ keybuf, long long B, int n, int W64, uint8_t *dag_bad, Header *hdr) {
    const int lane = (int)(threadIdx.x & 31);
    const long long b = (long long)blockIdx.x * ACYC_WARPS + (threadIdx.x >> 5);
    if (b >= B) return;
    bool bad = dag_bad[b] != 0;   // uniform over the warp
    if (!bad) {
        const int Wk = W64 + 1;
        const u64 *keys = keybuf + b * (long long)n * Wk;
        u64 pm[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = lane + 32 * j;
            pm[j][0] = i < n ? keys[(long long)i * Wk + 1] : 0ull;
            pm[j][1] = (i < n && W64 > 1) ? keys[(long long)i * Wk + 2] : 0ull;
        }
        u32 alive[4];   // 32 vertices per word, the same in every lane
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int bits = min(32, max(0, n - 32 * j));
            alive[j] = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
        }
        while (true) {
            const u64 a0 = (u64)alive[0] | ((u64)alive[1] << 32), a1 = (u64)alive[2] | ((u64)alive[3] << 32);
            u32 any_rm = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (32 * j >= n) break;
                const bool mine = (alive[j] >> lane) & 1u;
                const bool blocked = ((pm[j][0] & a0) | (pm[j][1] & a1)) != 0ull;
                const u32 rm = __ballot_sync(0xffffffffu, mine && !blocked);
                alive[j] &= ~rm;
                any_rm |= rm;
            }
            if ((alive[0] | alive[1] | alive[2] | alive[3]) == 0u) break;
            if (!any_rm) { bad = true; break; }
        }
    }
    if (lane == 0 && bad) {
        dag_bad[b] = 1;
        atomicAdd(&hdr->n_invalid, 1u);
    }
}

// Networks of more than 128 variables: one 256-thread block per DAG.  The alive set lives in shared
// memory as 32-bit words; warp w tests the 32 vertices of words w, w + 8, ... and writes a word's
// successor from one ballot, so a round needs no atomics (the first version cleared bits with 64-bit
// shared-memory atomicAnd, a CAS loop: 45 us per launch of 64 DAGs of 413 vertices, most of it the
// contention of the first rounds).  The DAG's parent masks are staged in shared memory when they fit.
__global__ void __launch_bounds__(ACYC_WIDE_THREADS)
k_acyclic_wide(const u64 *__restrict__ keybuf, long long B, int n, int W64, uint8_t *dag_bad, Header *hdr, int staged) {
    extern __shared__ u64 s_masks[];   // [n][W64] when staged
    __shared__ __align__(8) u32 s_alive[2][2 * W64MAX];
    const int lane = (int)(threadIdx.x & 31), warp = (int)(threadIdx.x >> 5);
    const long long b = blockIdx.x;
    if (b >= B) return;
    const int Wk = W64 + 1;
    const int nwords = (n + 31) >> 5;
    bool bad = dag_bad[b] != 0;        // uniform over the block
    if (!bad) {
        const u64 *keys = keybuf + b * (long long)n * Wk;
        for (int w = (int)threadIdx.x; w < 2 * W64; w += ACYC_WIDE_THREADS) {
            const int bits = min(32, max(0, n - w * 32));
            const u32 m = bits >= 32 ? 0xffffffffu : ((1u << bits) - 1u);
            s_alive[0][w] = m;
            s_alive[1][w] = m;
        }
        if (staged)
            for (int x = (int)threadIdx.x; x < n * W64; x += ACYC_WIDE_THREADS) s_masks[x] = keys[(long long)(x / W64) * Wk + 1 + x % W64];
        __syncthreads();
        int cur = 0;
        while (true) {
            const u32 *alive = s_alive[cur];
            const u64 *alive64 = reinterpret_cast<const u64 *>(alive);
            u32 *next = s_alive[cur ^ 1];
            u32 any_rm = 0, any_left = 0;
            for (int wd = warp; wd < nwords; wd += ACYC_WIDE_THREADS / 32) {
                const int i = wd * 32 + lane;
                const u32 aw = alive[wd];
                bool removable = false;
                if (i < n && ((aw >> lane) & 1u)) {
                    const u64 *pm = staged ? s_masks + (long long)i * W64 : keys + (long long)i * Wk + 1;
                    bool blocked = false;
                    for (int w = 0; w < W64; ++w) blocked = blocked || ((pm[w] & alive64[w]) != 0);
                    removable = !blocked;
                }
                const u32 rm = __ballot_sync(0xffffffffu, removable);
                const u32 left = aw & ~rm;
                if (lane == 0) next[wd] = left;
                any_rm |= rm;
                any_left |= left;
            }
            const int changed = __syncthreads_or((int)(any_rm != 0u));   // also publishes `next`
            const int any = __syncthreads_or((int)(any_left != 0u));
            cur ^= 1;
            if (!any) break;
            if (!changed) { bad = true; break; }
        }
    }
    if (threadIdx.x == 0 && bad) {
        dag_bad[b] = 1;
        atomicAdd(&hdr->n_invalid, 1u);
    }
}

// Without the cycle check the rejected DAGs (self loop, bad index) still have to be counted.
__global__ void k_count_bad(const uint8_t *dag_bad, long long B, Header *hdr) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && dag_bad[b]) atomicAdd(&hdr->n_invalid, 1u);
}

// ------------------------------------------------------------------------ cache: probe
// inst[t] >= 0: family id (cache hit);  -1: instance of a rejected DAG;  <= -2: slot -2-inst
// holds a PENDING entry of this batch.  The PENDING entry keeps the *smallest* instance index
// among duplicates, so the owner — and with it every family id — does not depend on thread
// timing (row-sharded ranks must agree on ids and table offsets).
__global__ void k_probe(const u64 *__restrict__ keybuf, int Wk, long long T, int n_per_dag,
                        const uint8_t *__restrict__ dag_bad, u32 *table, u32 mask,
                        const u64 *__restrict__ regkeys, int *inst) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    if (n_per_dag && dag_bad[t / n_per_dag]) { inst[t] = -1; return; }
    const u64 *key = keybuf + t * Wk;
    u32 s = (u32)hash_key(key, Wk) & mask;
    const u32 mine = ENT_PENDING | (u32)t;
    while (true) {
        u32 e = *((volatile u32 *)(table + s));
        if (e == ENT_EMPTY) {
            u32 old = atomicCAS(table + s, ENT_EMPTY, mine);
            if (old == ENT_EMPTY) { inst[t] = -2 - (int)s; return; }
            e = old;
        }
        if (e & ENT_PENDING) {
            long long t2 = (long long)(e & ~ENT_PENDING);
            if (keys_equal(key, keybuf + t2 * Wk, Wk)) {
                // the entry only ever decreases: a duplicate that cannot lower it skips the atomic (sachs: 1.1 M
                // instances of 9 k families, most of them later in the batch than the current owner)
                if (mine < e) atomicMin(table + s, mine);
                inst[t] = -2 - (int)s;
                return;
            }
        } else {
            long long id = (long long)e - 1;
            if (keys_equal(key, regkeys + id * Wk, Wk)) { inst[t] = (int)id; return; }
        }
        s = (s + 1) & mask;
    }
}

__global__ void k_owner_flags(const int *__restrict__ inst, const u32 *__restrict__ table, long long T,
                              u32 *flag) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int v = inst[t];
    flag[t] = (v <= -2 && table[-2 - v] == (ENT_PENDING | (u32)t)) ? 1u : 0u;
}

// ------------------------------------------------------------------------------- scan
// Exclusive scan, three launches; thread t of a block owns ITEMS consecutive elements.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_CHUNK = SCAN_THREADS * SCAN_ITEMS;

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partial(const TI *__restrict__ in, long long n, TO *bsum) {
    __shared__ TO sh[SCAN_THREADS / 32];
    long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_ITEMS;
    TO s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) s += (TO)in[base + i];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        TO tot = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) tot += sh[w];
        bsum[blockIdx.x] = tot;
    }
}

template <typename TO>
__global__ void __launch_bounds__(1024) k_scan_bsums(TO *bsum, int nb, TO *total) {
    __shared__ TO sh[1024];
    __shared__ TO carry;
    int tid = threadIdx.x;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        int i = base + tid;
        TO v = (i < nb) ? bsum[i] : (TO)0;
        sh[tid] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            TO x = (tid >= o) ? sh[tid - o] : (TO)0;
            __syncthreads();
            sh[tid] += x;
            __syncthreads();
        }
        TO incl = sh[tid], c = carry;
        __syncthreads();
        if (i < nb) bsum[i] = c + incl - v;
        if (tid == 1023) carry = c + incl;
        __syncthreads();
    }
    if (tid == 0) *total = carry;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const TI *__restrict__ in, long long n,
                                                             const TO *__restrict__ bsum, TO *out) {
    __shared__ TO sh[SCAN_THREADS];
    long long base = (long long)blockIdx.x * SCAN_CHUNK + (long long)threadIdx.x * SCAN_ITEMS;
    TO loc[SCAN_ITEMS];
    TO s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        loc[i] = (base + i < n) ? (TO)in[base + i] : (TO)0;
        s += loc[i];
    }
    int tid = threadIdx.x;
    sh[tid] = s;
    __syncthreads();
    for (int o = 1; o < SCAN_THREADS; o <<= 1) {
        TO x = (tid >= o) ? sh[tid - o] : (TO)0;
        __syncthreads();
        sh[tid] += x;
        __syncthreads();
    }
    TO run = bsum[blockIdx.x] + sh[tid] - s;
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = run;
        run += loc[i];
    }
}

// --------------------------------------------------------------- describe a new family
// Decodes a key into the count-kernel class, appends the job to its class list and adds the
// family's algorithmic bytes to the header.  Families whose table would exceed MAX_CELLS set
// err bit 0 and get no job.
__device__ __forceinline__ void describe_family(const u64 *key, int W64, const int *__restrict__ card,
                                                long long N, u32 j, u32 max_jobs, Header *hdr,
                                                u32 *cells_arr, int *class_jobs) {
    int node = (int)key[0];
    u64 r = (u64)card[node];
    u64 q = 1;
    int k = 0;
    u32 rad0 = 1;   // states of the first parent (the most significant digit of the cell index)
    bool over = false;
    for (int w = 0; w < W64; ++w) {
        u64 m = key[1 + w];
        while (m) {
            int b = __ffsll((long long)m) - 1;
            m &= m - 1;
            u64 c = (u64)card[w * 64 + b];
            if (c > 1) {
                if (k == 0) rad0 = (u32)c;
                ++k;
                q *= c;
                if (q > MAX_CELLS) { over = true; q = MAX_CELLS + 1; }
            }
        }
    }
    u64 cells = q * r;
    if (over || cells > MAX_CELLS) {
        atomicOr(&hdr->err, 1u);
        cells_arr[j] = 0;
        return;
    }
    cells_arr[j] = (u32)cells;
    atomicAdd(&hdr->cells_all, cells);
    int cls = cells <= CLASS0_CELLS ? 0 : cells <= CLASS1_CELLS ? 1 : cells <= CLASS2_CELLS ? 2 : 3;
    u32 pos = atomicAdd(&hdr->class_count[cls], 1u);
    class_jobs[(long long)cls * max_jobs + pos] = (int)j;
    atomicAdd(&hdr->alg_bytes[cls], (u64)(k + 1) * (u64)N + 4ull * cells);
    atomicAdd(&hdr->class_cells[cls], cells);
    atomicMax(&hdr->max_cells, (u32)cells);
    if (cls == 3) {
        const u32 p32 = range_plan((u32)cells, k, rad0, CLASS2_CELLS, true).passes;
        const u32 p16 = range_plan((u32)cells, k, rad0, k <= 6 ? 2u * CLASS2_CELLS : CLASS2_CELLS, false).passes;
        atomicMax(&hdr->max_passes3, p32);
        atomicMax(&hdr->max_passes3_u16, p16);
        atomicAdd(&hdr->sum_passes3, p32);
        atomicAdd(&hdr->sum_passes3_u16, p16);
    }
}

// Owners of PENDING entries take id = base + rank, publish their key in the registry and
// finalise the table entry.
__global__ void k_finalize(const u64 *__restrict__ keybuf, int Wk, long long T, int *inst,
                           const u32 *__restrict__ flag, const u32 *__restrict__ rank, long long base,
                           u64 *regkeys, u32 *table) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || !flag[t]) return;
    long long id = base + rank[t];
    const u64 *key = keybuf + t * Wk;
    for (int w = 0; w < Wk; ++w) regkeys[id * Wk + w] = key[w];
    table[-2 - inst[t]] = (u32)(id + 1);
    inst[t] = (int)id;
}

// Read-only lookup of a key among the finalised entries.
__device__ __forceinline__ long long cache_find(const u64 *key, int Wk, const u32 *__restrict__ table, u32 mask,
                                                const u64 *__restrict__ regkeys) {
    u32 s = (u32)hash_key(key, Wk) & mask;
    while (true) {
        u32 e = table[s];
        if (e == ENT_EMPTY) return -1;
        long long id = (long long)e - 1;
        if (keys_equal(key, regkeys + id * Wk, Wk)) return id;
        s = (s + 1) & mask;
    }
}

// Donor search.  A count table is the joint contingency table of the family's variables
// {node} + P; which of them is the child only fixes the axis order.  Counts are additive over a
// variable's states, so the table of family (y, Q) is the table of ANY family whose variable set
// contains {y} + Q, summed over the extra variables and re-ordered — no pass over the rows.
// The search is inverted: every new family G announces itself as a donor to all proper
// sub-variable-sets S of its own set V (every child designation c in S; read-only probes for
// the key (c, S - {c})); a hit that is new in this sub-batch records G with atomicMin on
// (joint cells of G, G's index), so the cheapest donor wins and the plan — like the ids — does
// not depend on thread timing.  |V| = 7 costs 7 * 2^6 = 448 probes; sets of more than
// ANNOUNCE_MAX_VARS variables only announce to subsets missing one or two variables.
constexpr int ANNOUNCE_MAX_VARS = 9;

__device__ __forceinline__ void announce_to(const u64 *sub, int W64, u64 pack, long long base, const u32 *__restrict__ table,
                                            u32 mask, const u64 *__restrict__ regkeys, u64 *best) {
    u64 key[W64MAX + 1];
    int Wk = W64 + 1;
    for (int w = 0; w < W64; ++w) {
        u64 bits = sub[w];
        while (bits) {
            int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            int c = w * 64 + b;                       // child designation
            key[0] = (u64)c;
            for (int v = 0; v < W64; ++v) key[1 + v] = sub[v];
            key[1 + w] &= ~(1ull << b);
            long long id = cache_find(key, Wk, table, mask, regkeys);
            if (id >= base) atomicMin(best + (id - base), pack);
        }
    }
}

// One warp per donor: the lanes share out its subsets (the probes are independent, the minimum
// is order-free).  world > 1 (family sharding): rank r announces only for the donors g = r,
// r + world, ...; the per-rank minima are then combined with ncclAllReduce(uint64, min).
__global__ void k_announce(const u64 *__restrict__ regkeys, int W64, long long base, const Header *hdr,
                           const u32 *__restrict__ table, u32 mask, const int *__restrict__ card, u64 *best,
                           int rank, int world) {
    const int lane = threadIdx.x & 31;
    long long g = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * world + rank;
    if (g >= hdr->f_new) return;
    int Wk = W64 + 1;
    const u64 *mine = regkeys + (base + g) * Wk;
    u64 vset[W64MAX], sub[W64MAX];
    int vars[DERIVE_LEVELS + 1];
    int node = (int)mine[0];
    int m = 0;
    u64 cells = 1;
    for (int w = 0; w < W64; ++w) vset[w] = mine[1 + w];
    vset[node >> 6] |= 1ull << (node & 63);
    for (int w = 0; w < W64; ++w) {
        u64 bits = vset[w];
        while (bits) {
            int b = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (m <= DERIVE_LEVELS) vars[m] = w * 64 + b;
            ++m;
            cells *= (u64)card[w * 64 + b];
            if (cells > MAX_CELLS) return;            // this family errors out anyway
        }
    }
    if (m < 2 || m > DERIVE_LEVELS) return;
    const u64 pack = (cells << 32) | (u64)g;
    if (m <= ANNOUNCE_MAX_VARS) {
        for (u32 pick = 1 + lane; pick + 1 < (1u << m); pick += 32) {   // every non-empty proper subset
            for (int w = 0; w < W64; ++w) sub[w] = 0;
            for (int i = 0; i < m; ++i)
                if (pick >> i & 1u) sub[vars[i] >> 6] |= 1ull << (vars[i] & 63);
            announce_to(sub, W64, pack, base, table, mask, regkeys, best);
        }
    } else {
        for (int i = lane; i < m; i += 32) {
            for (int w = 0; w < W64; ++w) sub[w] = vset[w];
            sub[vars[i] >> 6] &= ~(1ull << (vars[i] & 63));
            announce_to(sub, W64, pack, base, table, mask, regkeys, best);
            for (int i2 = i + 1; i2 < m && m > 2; ++i2) {
                sub[vars[i2] >> 6] &= ~(1ull << (vars[i2] & 63));
                announce_to(sub, W64, pack, base, table, mask, regkeys, best);
                sub[vars[i2] >> 6] |= 1ull << (vars[i2] & 63);
            }
        }
    }
}

__global__ void k_pick_donor(const u64 *__restrict__ best, const Header *hdr, int enable, int *donor) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hdr->f_new) return;
    u64 b = best[j];
    donor[j] = (enable && b != ~0ull) ? (int)(b & 0xffffffffull) : -1;
}

// Describe the new families: counted ones get a count job, derived ones go to the derive list.
// donor == nullptr: derivation is off, every new family is counted.
// world > 1: a family belongs to the rank that owns the root of its donor chain, so a derived
// family and every table it is derived from live on the same rank.  Family sharding (filter):
// families of other ranks get no job and no table here.  Row sharding with the fused
// reduce-scatter (owner_out): every rank counts every family on its rows, the owner reduces.
__global__ void k_describe_new(const u64 *__restrict__ regkeys, int W64, long long base, const int *__restrict__ card,
                               long long N, u32 max_jobs, Header *hdr, const int *__restrict__ donor,
                               u32 *cells_arr, int *class_jobs, int *derived_list, int rank, int world,
                               int filter, int *owner_out) {
    long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hdr->f_new) return;
    if (world > 1) {
        long long root = j;
        while (donor && donor[root] >= 0) root = donor[root];
        const int own = (int)(mix64((u64)root) % (u64)world);
        if (owner_out) owner_out[j] = own;
        if (filter && own != rank) {   // family sharding: another rank counts this family
            cells_arr[j] = 0;
            return;
        }
    } else if (owner_out) {
        owner_out[j] = 0;
    }
    const u64 *key = regkeys + (base + j) * (W64 + 1);
    if (!donor || donor[j] < 0) {
        describe_family(key, W64, card, N, (u32)j, max_jobs, hdr, cells_arr, class_jobs);
        return;
    }
    // derived: its table is smaller than its donor's, which passed the size check
    u64 cells = (u64)card[(int)key[0]];
    int pc = 0;
    for (int w = 0; w < W64; ++w) {
        u64 m = key[1 + w];
        pc += __popcll(m);
        while (m) {
            int b = __ffsll((long long)m) - 1;
            m &= m - 1;
            cells *= (u64)card[w * 64 + b];
            if (cells > MAX_CELLS) cells = MAX_CELLS;
        }
    }
    cells_arr[j] = (u32)cells;
    atomicAdd(&hdr->cells_all, cells);
    derived_list[atomicAdd(&hdr->n_derived, 1u)] = (int)j;
    atomicAdd(&hdr->lvl_count[pc], 1u);
}

// Group the derived families by level (number of parents), highest level first, so that each
// k_derive launch gets exactly its own families.  Order inside a level does not matter.
__global__ void k_group_derived(const u64 *__restrict__ regkeys, int W64, long long base, Header *hdr,
                                const int *__restrict__ derived_list, int *derived_by_level) {
    u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hdr->n_derived) return;
    const int j = derived_list[idx];
    const u64 *key = regkeys + (base + j) * (W64 + 1);
    int pc = 0;
    for (int w = 0; w < W64; ++w) pc += __popcll(key[1 + w]);
    u32 start = 0;
    for (int l = DERIVE_LEVELS - 1; l > pc; --l) start += hdr->lvl_count[l];
    derived_by_level[start + atomicAdd(&hdr->lvl_cursor[pc], 1u)] = j;
}

// Cache-bypassing path (bic_count_families): every listed family is job t.
__global__ void k_describe_direct(const u64 *__restrict__ keybuf, int W64, long long T,
                                  const int *__restrict__ card, long long N, u32 max_jobs, Header *hdr,
                                  u32 *cells_arr, int *class_jobs) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    describe_family(keybuf + t * (W64 + 1), W64, card, N, (u32)t, max_jobs, hdr, cells_arr, class_jobs);
}

// Experiment (BIC_SORT_JOBS=1, off by default): order the jobs of one class by their number of
// parents, most first (counting sort in shared memory, one block), so that CTAs resident together
// run the same specialisation of the ~300 KB count kernel (instruction cache) and take about as
// long as their neighbours.  Measured 2.2 % SLOWER on the alarm-shaped step (58.4 vs 57.2 ms):
// in the arbitrary order small replicated tables (ALU-bound) and large ones (bound by the
// shared-memory atomic pipe) share an SM and load the two pipes more evenly.  No result depends
// on the job order.
__global__ void __launch_bounds__(1024) k_order_jobs(const int *__restrict__ in, int cnt, const u64 *__restrict__ keys,
                                                     long long key_base, int W64, int *out) {
    __shared__ u32 s_hist[16], s_start[16];
    if (threadIdx.x < 16) s_hist[threadIdx.x] = 0;
    __syncthreads();
    auto bucket = [&](int j) {
        const u64 *key = keys + (key_base + j) * (long long)(W64 + 1);
        int pc = 0;
        for (int w = 0; w < W64; ++w) pc += __popcll(key[1 + w]);
        return min(pc, 15);
    };
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) atomicAdd(&s_hist[bucket(in[i])], 1u);
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 pos = 0;
        for (int b = 15; b >= 0; --b) { s_start[b] = pos; pos += s_hist[b]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int j = in[i];
        out[atomicAdd(&s_start[bucket(j)], 1u)] = j;
    }
}

// Which jobs need their table in HBM: S > 1 slices, class 3, or the caller wants the counts.
struct NeedArgs { int S[NCLASS]; int all; };
__global__ void k_table_need(const u32 *__restrict__ cells_arr, u32 njobs, NeedArgs a, u32 *need) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= njobs) return;
    u32 cells = cells_arr[j];
    int cls = cells <= CLASS0_CELLS ? 0 : cells <= CLASS1_CELLS ? 1 : cells <= CLASS2_CELLS ? 2 : 3;
    need[j] = (a.all || cls == 3 || a.S[cls] > 1) ? cells : 0u;
}

// -------------------------------------------------------------------- rehash on growth
__global__ void k_rehash(const u64 *__restrict__ regkeys, int Wk, long long count, u32 *table, u32 mask) {
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    u32 s = (u32)hash_key(regkeys + id * Wk, Wk) & mask;
    while (atomicCAS(table + s, ENT_EMPTY, (u32)(id + 1)) != ENT_EMPTY) s = (s + 1) & mask;
}

// ------------------------------------------------------------------------------ gather
__device__ __forceinline__ long long resolve_id(int v, const u32 *__restrict__ table) {
    return v >= 0 ? (long long)v : (long long)table[-2 - v] - 1;
}

// Per-DAG sum of its n family terms, in variable order (matches the oracle's summation order).
__global__ void k_gather_dags(const int *__restrict__ inst, const u32 *__restrict__ table, long long B, int n,
                              const uint8_t *__restrict__ dag_bad, const double *__restrict__ ll,
                              const double *__restrict__ np, double pen, double *out) {
    long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (dag_bad[b]) { out[b] = __longlong_as_double(0x7ff8000000000000LL); return; }
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        long long id = resolve_id(inst[b * n + i], table);
        s += ll[id] - pen * np[id];
    }
    out[b] = s;
}

// The same sum for wide networks (n >= 64): one warp per DAG.  The lanes fetch 32 terms at a time,
// lane 0 adds them in variable order, so the bits equal the one-thread version's.
__global__ void k_gather_dags_warp(const int *__restrict__ inst, const u32 *__restrict__ table, long long B, int n,
                                   const uint8_t *__restrict__ dag_bad, const double *__restrict__ ll,
                                   const double *__restrict__ np, double pen, double *out) {
    const int lane = threadIdx.x & 31;
    long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    if (dag_bad[b]) { if (lane == 0) out[b] = __longlong_as_double(0x7ff8000000000000LL); return; }
    double s = 0.0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        double term = 0.0;
        if (i0 + lane < n) {
            long long id = resolve_id(inst[b * n + i0 + lane], table);
            term = ll[id] - pen * np[id];
        }
        const int cnt = min(32, n - i0);
        for (int k = 0; k < cnt; ++k) s += __shfl_sync(0xffffffffu, term, k);
    }
    if (lane == 0) out[b] = s;
}

// Small warm batches in one launch (n <= 64, adjacency input): keys, self-loop and cycle check,
// read-only cache lookup and the per-DAG sum, one warp per DAG.  The reference's own usage is one
// score() call per DAG in a Python loop (src/predictors/utils.py:22-24); with a warm cache the
// general pipeline (~25 launches, two host synchronisations) costs ~80 us per call, this kernel
// one launch and one synchronisation.  Any family that is not cached yet sets bit 3 of hdr->err
// and the host falls back to the general pipeline (which inserts and counts it).  Sums run in
// variable order with the same expression as k_gather_dags: identical bits.
constexpr int SMALL_WARPS = 4;
__global__ void __launch_bounds__(SMALL_WARPS * 32)
k_score_small(const uint8_t *__restrict__ adj, long long B, int n, const u32 *__restrict__ table, u32 mask,
              const u64 *__restrict__ regkeys, const double *__restrict__ ll, const double *__restrict__ np,
              double pen, int check_cycles, double *out, Header *hdr) {
    __shared__ u64 s_pm[SMALL_WARPS][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * SMALL_WARPS + warp;
    if (b >= B) return;
    const uint8_t *a = adj + b * (long long)n * n;
    u64 *pm = s_pm[warp];
    bool bad = false;
    for (int i = lane; i < n; i += 32) {
        u64 m = 0;
        for (int p = 0; p < n; ++p)
            if (a[(long long)p * n + i]) m |= 1ull << p;
        pm[i] = m;
        bad = bad || ((m >> i) & 1ull);   // self loop
    }
    __syncwarp();
    bad = __any_sync(0xffffffffu, bad);
    if (!bad && check_cycles) {   // peel parentless vertices; alive is uniform over the warp
        u64 alive = n == 64 ? ~0ull : ((1ull << n) - 1ull);
        while (alive) {
            const bool f0 = lane < n && ((alive >> lane) & 1ull) && (pm[lane] & alive) == 0;
            const bool f1 = lane + 32 < n && ((alive >> (lane + 32)) & 1ull) && (pm[lane + 32] & alive) == 0;
            const u64 rem = (u64)__ballot_sync(0xffffffffu, f0) | ((u64)__ballot_sync(0xffffffffu, f1) << 32);
            if (!rem) { bad = true; break; }
            alive &= ~rem;
        }
    }
    if (bad) {
        if (lane == 0) {
            out[b] = __longlong_as_double(0x7ff8000000000000LL);
            atomicAdd(&hdr->n_invalid, 1u);
        }
        return;
    }
    double term[2] = {0.0, 0.0};
    bool miss = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < n) {
            const u64 key[2] = {(u64)i, pm[i]};
            const long long id = cache_find(key, 2, table, mask, regkeys);
            if (id < 0) miss = true;
            else term[h] = ll[id] - pen * np[id];
        }
    }
    if (__any_sync(0xffffffffu, miss)) {
        if (lane == 0) atomicOr(&hdr->err, 8u);
        return;
    }
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += __shfl_sync(0xffffffffu, i < 32 ? term[0] : term[1], i & 31);
    if (lane == 0) out[b] = s;
}

__global__ void k_gather_fams(const int *__restrict__ inst, const u32 *__restrict__ table, long long T,
                              const double *__restrict__ ll, const double *__restrict__ np, double pen,
                              double *out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    long long id = resolve_id(inst[t], table);
    out[t] = ll[id] - pen * np[id];
}

// Offsets of the new families inside a slot of their owner's exchange buffer (row-sharded runs
// with the fused reduce-scatter): block w walks all jobs and packs those of owner w back to back
// (derived families get no space: only counted tables travel).  hdr->owned_max = the largest
// total, which must fit a slot.  Deterministic: every rank computes the same offsets.
__global__ void __launch_bounds__(1024) k_owner_offsets(const u32 *__restrict__ cells_arr, const int *__restrict__ owner,
                                                        const int *__restrict__ donor, Header *hdr, u64 *xoff) {
    __shared__ u64 sh[1024];
    __shared__ u64 carry;
    const int w = blockIdx.x, tid = threadIdx.x;
    const u32 njobs = hdr->f_new;
    if (tid == 0) carry = 0;
    __syncthreads();
    for (u32 base = 0; base < njobs; base += 1024) {
        const u32 i = base + tid;
        const bool mine = i < njobs && owner[i] == w && !(donor && donor[i] >= 0);
        const u64 v = mine ? (u64)((cells_arr[i] + 3u) & ~3u) : 0ull;   // 16-byte aligned tables
        sh[tid] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            u64 x = (tid >= o) ? sh[tid - o] : 0ull;
            __syncthreads();
            sh[tid] += x;
            __syncthreads();
        }
        const u64 incl = sh[tid], c = carry;
        __syncthreads();
        if (mine) xoff[i] = c + incl - v;
        if (tid == 1023) carry = c + incl;
        __syncthreads();
    }
    if (tid == 0) atomicMax(&hdr->owned_max, carry);
}

// ------------------------------------------------------------------- dataset validation
// 16 rows per load; a byte >= card is found with per-byte compares on the four words.
__global__ void k_validate(const uint8_t *__restrict__ data, long long N, long long stride, int n,
                           const int *__restrict__ card, u32 *bad) {
    int v = blockIdx.y;
    const u32 c = (u32)card[v];
    const uint4 *col = reinterpret_cast<const uint4 *>(data + (long long)v * stride);
    const long long nvec = (N + 15) >> 4;   // tail rows of the padded column hold state 0
    bool b = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const uint4 w = col[i];
        const u32 ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int s = 0; s < 4; ++s) b = b || (((ws[k] >> (8 * s)) & 0xffu) >= c);
    }
    if (b) atomicOr(bad, 1u);
}

// 64-bit content fingerprint of the dataset (cache checkpoints refuse another dataset of the same
// shape): order-independent sum of mixed (position, 16-row word) pairs.
__global__ void k_fingerprint(const uint8_t *__restrict__ data, long long N, long long stride, int n, u64 *out) {
    int v = blockIdx.y;
    const uint4 *col = reinterpret_cast<const uint4 *>(data + (long long)v * stride);
    const long long nvec = (N + 15) >> 4;
    u64 acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const uint4 w = col[i];
        const u64 pos = (u64)v * 0x9E3779B97F4A7C15ULL + (u64)i;
        acc += mix64(pos ^ mix64(((u64)w.x << 32 | w.y) + 0x632BE59BD9B4E019ULL * (((u64)w.z << 32) | w.w)));
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

// Imported cache keys (bic_cache_import): node < n, parent bits < n, node not its own parent.
__global__ void k_check_keys(const u64 *__restrict__ keys, int W64, long long count, int n, u32 *bad) {
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    const u64 *key = keys + id * (W64 + 1);
    bool b = key[0] >= (u64)n;
    for (int w = 0; w < W64; ++w) {
        const int bits = min(64, max(0, n - w * 64));
        const u64 valid = bits >= 64 ? ~0ull : ((1ull << bits) - 1ull);
        b = b || (key[1 + w] & ~valid);
    }
    if (!b) b = (key[1 + (key[0] >> 6)] >> (key[0] & 63)) & 1ull;
    if (b) atomicOr(bad, 1u);
}

// After k_rehash of imported keys: a key present twice sits in two slots; the second lookup of its
// own key finds the other id.
__global__ void k_check_duplicates(const u64 *__restrict__ regkeys, int Wk, long long count, const u32 *__restrict__ table,
                                   u32 mask, u32 *bad) {
    long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= count) return;
    if (cache_find(regkeys + id * Wk, Wk, table, mask, regkeys) != id) atomicOr(bad, 2u);
}

// ------------------------------------------------------------ 2-bit shadow copy of the dataset
// Columns with at most 4 states also get a packed copy, 4 rows per byte (row p of a group of 16
// in bits 2p..2p+1 of a 32-bit word).  The count kernel streams it instead of the uint8 column
// when every column of a family qualifies: a quarter of the load traffic through the L1TEX data
// pipe, which is what limits the kernel.  One thread packs 16 rows.
__global__ void k_pack2(const uint8_t *__restrict__ data, long long stride, int n, const int *__restrict__ card,
                        uint8_t *data2, long long stride2) {
    int v = blockIdx.y;
    if (card[v] > 4) return;
    long long nvec = stride >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(data + (long long)v * stride);
    u32 *dst = reinterpret_cast<u32 *>(data2 + (long long)v * stride2);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        uint4 w = src[i];
        const u32 ws[4] = {w.x, w.y, w.z, w.w};
        u32 out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u32 x = ws[k];
            u32 b = (x & 3u) | (((x >> 8) & 3u) << 2) | (((x >> 16) & 3u) << 4) | (((x >> 24) & 3u) << 6);
            out |= b << (8 * k);
        }
        dst[i] = out;
    }
}

}  // namespace bic
