/* bicgpu.h — C ABI of libbicgpu.so, the B200 (sm_100a) BIC structure scorer.
 *
 * This is the drop-in boundary for the reference's score path
 *     src/problem/bn/bnlearn.py:27-61          BNLearnWrapper.score()
 *     src/problem/bn/bnlearn_scripts/bnlearn_score.R:7-40   (Rscript child process)
 * The reference crosses a *process* boundary per DAG (bnlearn.py:46-54 spawns Rscript and
 * parses one float from stdout).  A maintainer replaces that subprocess call with the entry
 * points below (ctypes binding shown in INTEGRATION.md).  Plain pointers and sizes only; no
 * torch / C++ types; no exception crosses the boundary; every function returns BIC_OK (0) or a
 * negative bic_status and leaves a message retrievable with bic_last_error().
 *
 * Conventions (identical to oracle/bic_oracle.py, which restates bnlearn's discrete BIC):
 *   - dataset: column-major uint8 state codes, codes[v * stride + row], variable v = v-th
 *     data column (bnlearn_score.R:29), code < card[v];
 *   - adjacency: uint8 [B][n][n], adj[b][p][c] != 0  <=>  edge p -> c (row = parent,
 *     bnlearn.py:44, bnlearn_score.R:35);
 *   - count table of family (node i, parents P sorted ascending, first most significant):
 *     cell = j * r_i + x_i,  j = ((x_p1 * r_p2 + x_p2) * r_p3 + ...), int32 counters;
 *   - score = sum_i [ sum_{jk: N_ijk>0} N_ijk ln(N_ijk / N_ij)  -  pen * (r_i - 1) * q_i ],
 *     pen = 0.5 ln N (bic), 1 (aic), 0 (loglik); q_i over *declared* cardinalities;
 *   - bde / k2: sum_i sum_j [ lgamma(a_ij) - lgamma(a_ij + N_ij) + sum_k ( lgamma(a_ijk + N_ijk)
 *     - lgamma(a_ijk) ) ],  a_ijk = iss / (q_i r_i) (bde) or 1 (k2),  a_ij = r_i a_ijk.
 *
 * Pointer arguments are HOST pointers unless BIC_FLAG_DEVICE_PTRS is passed, in which case
 * every array argument of that call (inputs and outputs) is a device pointer on the context's
 * GPU.  Pointers are borrowed for the duration of the call.  Calls on one context are
 * serialised; use one context per GPU (one process per GPU for multi-GPU runs).
 *
 * Stream ordering of device pointers.  The library launches on the context's own non-blocking
 * stream (or the one given to bic_set_stream) and every scoring call returns only after that
 * stream has drained, so OUTPUTS are complete on return.  INPUTS are the caller's business: work
 * that produces a device input on another stream (a decoder, a dtype conversion, an all-gather)
 * is NOT ordered before the library's kernels unless the caller either synchronises that stream,
 * hands it over with bic_set_stream, or calls bic_wait_stream(ctx, producer_stream) right before
 * the scoring call (an event on the producer stream that the context's stream waits on; no host
 * synchronisation).  The Python host layer does the latter for every CUDA tensor it is given.
 */
#ifndef BICGPU_H
#define BICGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BICGPU_VERSION 200

typedef struct bic_ctx bic_ctx;

typedef enum {
    BIC_OK = 0,
    BIC_ERR_CUDA = -1,            /* a CUDA runtime call or kernel failed                    */
    BIC_ERR_ARG = -2,             /* bad argument (null pointer, n out of range, ...)        */
    BIC_ERR_NO_DATASET = -3,      /* scoring call before bic_set_dataset                     */
    BIC_ERR_TABLE_TOO_LARGE = -4, /* a family's q*r exceeds the count-table limit            */
    BIC_ERR_OOM = -5,             /* device allocation failed                                */
    BIC_ERR_NCCL = -6,            /* NCCL missing or a collective failed                     */
    BIC_ERR_BAD_CODE = -7,        /* dataset holds a state code >= its declared cardinality  */
    BIC_ERR_BAD_FAMILY = -8       /* a parent index is out of range or equals the node       */
} bic_status;

/* bnlearn score(type = ...) names: "bic", "loglik", "aic", "bde" (BDeu with imaginary sample size
 * iss, see bic_set_iss), "k2".  The reference forwards the string unchanged (bnlearn.py:50,
 * bnlearn_score.R:38); only "bic" is pinned by its tests. */
typedef enum { BIC_METRIC_BIC = 0, BIC_METRIC_LOGLIK = 1, BIC_METRIC_AIC = 2, BIC_METRIC_BDE = 3, BIC_METRIC_K2 = 4 } bic_metric;

enum {
    BIC_FLAG_DEVICE_PTRS = 1,    /* array arguments are device pointers                       */
    BIC_FLAG_NO_CYCLE_CHECK = 2, /* skip the acyclicity check (bnlearn_score.R:35 does check) */
    BIC_FLAG_NO_CACHE = 4,       /* clear the family-score cache before this call             */
    BIC_FLAG_NO_DERIVE = 8,      /* count every new family from the rows, even when its table  *
                                  * could be marginalised from a counted superset family       */
    BIC_FLAG_LOCAL_BATCH = 16    /* bic_score_dags_*, family sharding only: the arrays hold THIS *
                                  * rank's B DAGs (same B on every rank); the library all-gathers *
                                  * the family keys over NVLink, deduplicates the union on every  *
                                  * rank and returns the B local scores                           */
};

/* ---- lifetime ------------------------------------------------------------------------- */
int bic_version(void);
/* Build provenance: ABI version, target architecture, compiler and build time of the loaded library
 * (bench.py prints it into its JSON line, so a measurement names the binary it was taken on). */
const char *bic_build_info(void);
/* Replaces: process start-up of the Rscript child (bnlearn.py:46-54). */
int bic_create(bic_ctx **out, int device);
int bic_destroy(bic_ctx *ctx);
/* Message of the last failing call on ctx (ctx == NULL: last bic_create failure). */
const char *bic_last_error(const bic_ctx *ctx);
/* Launch on a caller-owned cudaStream_t (e.g. torch's current stream) instead of the
 * context's own stream.  NULL restores the own stream. */
int bic_set_stream(bic_ctx *ctx, void *cuda_stream);
int bic_sync(bic_ctx *ctx);
/* Order everything queued so far on `producer_stream` (a cudaStream_t; NULL = the legacy default
 * stream) before the kernels of the following calls on this context.  No host synchronisation.
 * See "Stream ordering of device pointers" above. */
int bic_wait_stream(bic_ctx *ctx, void *producer_stream);
/* Imaginary sample size of the "bde" metric (bnlearn's iss argument; default 1).  Clears cached
 * bde terms. */
int bic_set_iss(bic_ctx *ctx, double iss);

/* ---- dataset ---------------------------------------------------------------------------
 * Replaces: data(list = dataset_name); dataset <- get(dataset_name)  (bnlearn_score.R:25-26).
 * codes: uint8 [n][stride] column-major (stride >= N), card: int32 [n] declared cardinalities
 * (1..255).  The library keeps its own padded copy in HBM and validates codes < card.
 * Clears the family-score cache. */
int bic_set_dataset(bic_ctx *ctx, const uint8_t *codes, int64_t N, int32_t n, int64_t stride,
                    const int32_t *card, int is_device);
/* 64-bit content fingerprint of the dataset held by the context (computed on the device; depends
 * on every code, N and n).  Cache checkpoints carry it so that a checkpoint made on other rows of
 * the same shape (a bootstrap resample, another row shard) is refused. */
int bic_dataset_fingerprint(bic_ctx *ctx, uint64_t *out);

/* ---- families --------------------------------------------------------------------------
 * Replaces: the per-node contingency counting inside bnlearn::score (call site
 * bnlearn_score.R:38).  Family f = (node[f], parents[parent_off[f] .. parent_off[f+1])).
 * counts_out receives the dense int32 table of family f at counts_off[f]; counts_off[f+1] -
 * counts_off[f] must equal q_f * r_f.  Bypasses the cache.  flags: BIC_FLAG_DEVICE_PTRS. */
int bic_count_families(bic_ctx *ctx, const int32_t *node, const int64_t *parent_off,
                       const int32_t *parents, int64_t F, const int64_t *counts_off,
                       int32_t *counts_out, int flags);
/* One decomposable score term per family (through the family-score cache). */
int bic_score_families(bic_ctx *ctx, const int32_t *node, const int64_t *parent_off,
                       const int32_t *parents, int64_t F, int metric, double *out, int flags);

/* ---- DAGs ------------------------------------------------------------------------------
 * Replaces: BNLearnWrapper.score() end to end (bnlearn.py:38-61 + bnlearn_score.R:7-40) for a
 * batch of B DAGs.  out[b] = score, or NaN when DAG b is cyclic / has a self loop (the
 * reference's R child exits non-zero there, bnlearn.py:56-57); *n_invalid (may be NULL)
 * receives how many were rejected. */
int bic_score_dags_adj(bic_ctx *ctx, const uint8_t *adj, int64_t B, int metric, double *out,
                       int64_t *n_invalid, int flags);
/* Same, parent lists in CSR: family (b, i) has parents[off[b*n+i] .. off[b*n+i+1]);
 * off has B*n + 1 entries.  For wide networks (n in the hundreds) where B*n*n bytes of
 * adjacency would dominate. */
int bic_score_dags_csr(bic_ctx *ctx, const int64_t *off, const int32_t *parents, int64_t B,
                       int metric, double *out, int64_t *n_invalid, int flags);
/* Same, the reference's candidate wire format (src/toolkit/labeled.py:116-154): per DAG n
 * vertex labels (BN variable index of vertex v) and n edge words, bit u of ebits[b][v] set
 * <=> edge vertex u -> vertex v (u < v); the relabel of bnlearn.py:38-42 runs on the GPU.
 * n <= 32.  DAGs whose labels are not a permutation of 0..n-1 are rejected like cyclic ones
 * (bnlearn.py:34-35 asserts).
 * With BIC_FLAG_LOCAL_BATCH (any of the three DAG entry points, family sharding): B, the arrays,
 * out[] and *n_invalid all refer to this rank's own DAGs. */
int bic_score_dags_wire(bic_ctx *ctx, const uint8_t *labels, const uint32_t *ebits, int64_t B,
                        int metric, double *out, int64_t *n_invalid, int flags);
/* The wire format for any n <= 1024, with the reference's own label type (l_i is uint16,
 * src/toolkit/labeled.py:118; e_i has i characters, :124): labels [B][n] uint16, ebits
 * [B][n][ewords] with ewords >= ceil(n / 32), bit (u % 32) of word (u / 32) of ebits[b][v] set
 * <=> edge vertex u -> vertex v (u < v; bits at u >= v are ignored).  A label >= n or a repeated
 * label rejects the DAG (bnlearn.py:34-35). */
int bic_score_dags_wire16(bic_ctx *ctx, const uint16_t *labels, const uint32_t *ebits, int32_t ewords,
                          int64_t B, int metric, double *out, int64_t *n_invalid, int flags);

/* ---- family-score cache ---------------------------------------------------------------- */
typedef struct {
    int64_t families;      /* distinct (node, parent-set) families held                     */
    int64_t capacity;      /* families the current allocation can hold                      */
    int64_t lookups;       /* family instances looked up since creation / last clear        */
    int64_t misses;        /* of those, how many had to be counted                          */
    int64_t bytes;         /* device bytes held by table + registry                         */
} bic_cache_stats_t;
int bic_cache_clear(bic_ctx *ctx);
int bic_cache_reserve(bic_ctx *ctx, int64_t families);
int bic_cache_stats(bic_ctx *ctx, bic_cache_stats_t *out);
/* Checkpoint / resume of a long search (the reference checkpoints only its VAE,
 * experiments/01_bn_asia/main.py:187-188; its scorer is stateless).  Export copies the registry
 * to HOST buffers: keys [families][1 + ceil(n/64)] (word 0 = node, then the parent bitmask),
 * terms[families] (log-likelihood, or the bde / k2 term) and nparams[families]; *kind receives
 * what the terms are (0 log-likelihood, 1 bde, 2 k2).  Pass capacity = 0 to only query
 * *families.  Import replaces the cache content; the caller is responsible for using it with
 * the same dataset (and iss). */
int bic_cache_export(bic_ctx *ctx, uint64_t *keys, double *terms, double *nparams, int64_t capacity,
                     int64_t *families, int32_t *kind);
int bic_cache_import(bic_ctx *ctx, const uint64_t *keys, const double *terms, const double *nparams,
                     int64_t families, int32_t kind);

/* ---- profiling (CUDA events on the launching stream, around the family-count kernels) --- */
typedef struct {
    double count_ms;       /* device time spent in family-count launches                    */
    int64_t count_launches;/* how many family-count kernels were launched                   */
    int64_t kernel_launches;/* every kernel this library launched                           */
    int64_t families_counted;
    int64_t rows_counted;  /* sum over counted families of N (this rank's rows)             */
    int64_t alg_bytes;     /* sum over counted families of (k+1)*N + 4*q*r                  */
    /* the same, split by count-kernel class (0..2: shared-memory tables of <= 2048 / 12288 /
     * 49152 cells, kernels k_count<256,false> / <512,false> / <1024,false>; 3: larger tables,
     * counted in passes over shared-memory sub-ranges, k_count<1024,false,true>, when the rows
     * dwarf the table, else straight into HBM with L2 atomics, k_count<256,true>) */
    double class_ms[4];
    int64_t class_launches[4];
    int64_t class_families[4];
    int64_t class_alg_bytes[4];
    int64_t families_derived; /* new families whose table was marginalised from a superset's  */
    /* row-sharded runs: the exchange step after the count kernels (barrier or ncclAllReduce of the
     * tables + the fp64 reduce from HBM), CUDA-event time; bytes this rank sent: count tables stored
     * into peers' exchange buffers (fused reduce-scatter) or the all-reduced payload (NCCL path) */
    double exchange_ms;
    int64_t exchange_bytes;
    int64_t exchange_fused;   /* exchange steps that took the fused reduce-scatter (peer stores) */
    int64_t exchange_nccl;    /* exchange steps that all-reduced the tables with NCCL            */
} bic_profile_t;
int bic_profile_enable(bic_ctx *ctx, int on);
int bic_profile_reset(bic_ctx *ctx);
int bic_profile_get(bic_ctx *ctx, bic_profile_t *out);

/* ---- launch planning (host arithmetic only; no GPU needed) --------------------------------
 * How one batch of new families is cut into count-kernel work: row slices per family for each
 * table-size class and whether class 3 (tables above one CTA's shared memory) is counted in
 * shared-memory sub-range passes or straight into HBM.  The library calls the same function
 * internally; it is exported so that the planning rules can be tested without a GPU.  The result
 * never changes a count or a score, only the time it takes. */
typedef struct {
    int32_t sm_count;          /* SMs of the device (B200: 148)                                 */
    int64_t N;                 /* rows on this GPU                                              */
    int32_t n;                 /* variables                                                     */
    int64_t max_cells;         /* largest q*r among the families                                */
    int32_t tables_in_hbm;     /* 1: every table is merged into HBM anyway (row-sharded, derived
                                * families present, bic_count_families)                         */
    int64_t class_count[4];    /* families per class (<= 2048 / 12288 / 49152 cells / larger)   */
    int64_t class_cells[4];    /* sum of q*r per class                                          */
    int64_t class_alg_bytes[4];/* sum of (k+1)*N + 4*q*r per class                              */
    int32_t all_packed;        /* 1: every column also has a 2-bit packed copy that the families *
                                * stream instead (a quarter of the bytes per row)               */
    int32_t passes3;           /* most sub-range passes a class-3 family of the batch needs      *
                                * (sub-ranges may follow the first parent's states);            *
                                * 0: ceil(max_cells / 49152)                                    */
    int32_t items3;            /* class-3 work items per row slice (sum of the families' passes); *
                                * 0: class_count[3] * passes                                    */
} bic_plan_in_t;
typedef struct {
    int32_t slices[4];         /* row slices per family, per class                              */
    int32_t ranged;            /* 1: class 3 in shared-memory sub-range passes                  */
    int32_t passes;            /* sub-range passes per (family, slice) when ranged, else 1      */
    int32_t cluster;           /* > 0: class 3 in one pass over thread-block clusters of this   *
                                * many CTAs that share the table in distributed shared memory   */
} bic_plan_out_t;
int bic_plan_slices(const bic_plan_in_t *in, bic_plan_out_t *out);
/* How one class-3 family (q*r = cells > 49152, k parents with more than one state, the first of
 * them with rad0 states) is cut into the sub-ranges a CTA counts in shared memory: cells per
 * sub-range, number of passes over the rows, and - when the sub-ranges follow the states of the
 * first parent - the states per pass (else 0).  counters16 != 0: the 16-bit-counter variant
 * (sub-ranges of twice the cells, k <= 6).  Host arithmetic shared with the kernels
 * (range_plan, csrc/common.cuh). */
int bic_range_plan(uint32_t cells, int32_t k, uint32_t rad0, int32_t counters16, uint32_t *span, uint32_t *passes,
                   uint32_t *states_per_pass);

/* ---- row sharding over several GPUs (one process per GPU) ------------------------------
 * Each rank holds N_rank rows of the same n columns.  After bic_comm_init every scoring call
 * must be made collectively with identical family / DAG arguments on every rank.  The partial
 * count tables are summed by a reduce-scatter fused into the count kernels: every family has an
 * owner rank, the CTA that completes a family's local table stores it into the owner's exchange
 * buffer with peer stores over NVLink (buffers mapped with CUDA IPC at the first call), a 4-byte
 * all-reduce orders the stores, the owner sums the partial tables inside its fp64 reduce and the
 * family terms travel with one ncclAllReduce(double).  Where peer mapping is unavailable, the
 * caller wants the tables themselves (bic_count_families) or a batch exceeds the exchange buffer,
 * the tables are summed with ncclAllReduce(uint32) instead.  Either way the counts are integer
 * sums (bit-exact for any number of ranks), every rank ends with identical bits and ln N uses
 * the global row count.  NCCL is dlopen()ed ("libnccl.so.2") on first use.
 * bic_comm_unique_id: rank 0 creates the 128-byte id and ships it to the others (e.g. with
 * torch.distributed.broadcast). */
int bic_comm_unique_id(uint8_t id_out[128]);
int bic_comm_init(bic_ctx *ctx, const uint8_t id[128], int rank, int world);
int bic_comm_destroy(bic_ctx *ctx);
/* What the communicator is used for.  BIC_SHARD_ROWS (default): each rank holds a slice of the
 * rows, count tables are all-reduced (above).  BIC_SHARD_FAMILIES: every rank holds the whole
 * dataset and is given the SAME (global) candidate batch; the batch is deduplicated identically
 * everywhere, each rank counts only the families whose donor tree it owns, and the family terms
 * are combined with ncclAllReduce(double, sum) — every other rank contributes exact zeros, so
 * the scores are bit-identical to a single-GPU run.  Clears the cache. */
enum { BIC_SHARD_ROWS = 0, BIC_SHARD_FAMILIES = 1 };
int bic_comm_mode(bic_ctx *ctx, int mode);

#ifdef __cplusplus
}
#endif
#endif /* BICGPU_H */
