/* CPU oracle for the BIC score path in plain C — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library (oracle/liboracle.so, built by oracle/Makefile).  The product library
 * (libbicgpu.so) never links, loads or calls it and has no CPU fallback.
 *
 * Restates the arithmetic reached from reference src/problem/bn/bnlearn.py:27-61 ->
 * src/problem/bn/bnlearn_scripts/bnlearn_score.R:38  score(net, dataset, type = "bic"),
 * i.e. the decomposable discrete BIC of the third-party R package bnlearn (un-vendored,
 * un-pinned; see oracle/bic_oracle.py for the formula, conventions and parity status:
 * asia/bic pinned by the reference's known answer + 1408 shipped values, everything else
 * "parity unpinned").  It is validated against oracle/bic_oracle.py in tests/test_oracle_c.py.
 *
 * Like the reference (bnlearn.py:46-54 spawns a fresh R process per DAG and bnlearn recounts
 * every family), oracle_score_dags_adj() recounts all n families of every DAG: no family cache.
 * Parallelism (OpenMP over (DAG, node) pairs) is the only liberty taken, so that the CPU
 * baseline can use every host core.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { ORACLE_BIC = 0, ORACLE_LOGLIK = 1, ORACLE_AIC = 2, ORACLE_BDE = 3, ORACLE_K2 = 4 };

static double g_iss = 1.0;   /* imaginary sample size of the BDeu metric (bnlearn iss) */
void oracle_set_iss(double iss) { g_iss = iss; }

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* q = product of declared parent cardinalities (bnlearn charges unobserved configs too). */
static int64_t family_q(const int32_t *card, const int32_t *parents, int k) {
    int64_t q = 1;
    for (int a = 0; a < k; ++a) q *= card[parents[a]];
    return q;
}

/* Dense table N_ijk, cell = j * r + x_i, j mixed-radix over parents in the given (ascending)
 * order with the first parent most significant.  counts must hold q * r zeroed int64. */
int oracle_family_counts(const uint8_t *codes, int64_t N, int64_t stride, const int32_t *card,
                         int32_t node, const int32_t *parents, int32_t k, int64_t *counts) {
    const uint8_t *xc = codes + (int64_t)node * stride;
    const int64_t r = card[node];
    for (int64_t row = 0; row < N; ++row) {
        int64_t j = 0;
        for (int a = 0; a < k; ++a)
            j = j * card[parents[a]] + codes[(int64_t)parents[a] * stride + row];
        counts[j * r + xc[row]] += 1;
    }
    return 0;
}

/* sum_{jk: N_ijk > 0} N_ijk * ln(N_ijk / N_ij) - penalty(metric). */
double oracle_score_counts(const int64_t *counts, int64_t q, int32_t r, int64_t N, int32_t metric) {
    if (metric == ORACLE_BDE || metric == ORACLE_K2) {
        /* bnlearn "bde" (BDeu, a_ijk = iss / (q r)) and "k2" (a_ijk = 1) */
        double a_ijk = metric == ORACLE_BDE ? g_iss / ((double)q * (double)r) : 1.0, a_ij = a_ijk * r, s = 0.0;
        for (int64_t j = 0; j < q; ++j) {
            int64_t nij = 0;
            for (int x = 0; x < r; ++x) nij += counts[j * r + x];
            if (!nij) continue;
            s += lgamma(a_ij) - lgamma(a_ij + (double)nij);
            for (int x = 0; x < r; ++x)
                if (counts[j * r + x]) s += lgamma(a_ijk + (double)counts[j * r + x]) - lgamma(a_ijk);
        }
        return s;
    }
    double ll = 0.0;
    for (int64_t j = 0; j < q; ++j) {
        int64_t nij = 0;
        for (int x = 0; x < r; ++x) nij += counts[j * r + x];
        if (!nij) continue;
        for (int x = 0; x < r; ++x) {
            int64_t c = counts[j * r + x];
            if (c) ll += (double)c * log((double)c / (double)nij);
        }
    }
    double nparams = (double)(r - 1) * (double)q;
    if (metric == ORACLE_BIC) return ll - (N > 0 ? 0.5 * log((double)N) * nparams : 0.0);
    if (metric == ORACLE_AIC) return ll - nparams;
    return ll;
}

double oracle_family_score(const uint8_t *codes, int64_t N, int64_t stride, const int32_t *card,
                           int32_t node, const int32_t *parents, int32_t k, int32_t metric) {
    int64_t q = family_q(card, parents, k);
    int32_t r = card[node];
    int64_t *counts = (int64_t *)calloc((size_t)(q * r), sizeof(int64_t));
    if (!counts) return NAN;
    oracle_family_counts(codes, N, stride, card, node, parents, k, counts);
    double s = oracle_score_counts(counts, q, r, N, metric);
    free(counts);
    return s;
}

/* Families given as CSR: node[f], parents[off[f] .. off[f+1]).  One score per family. */
int oracle_score_families(const uint8_t *codes, int64_t N, int64_t stride, const int32_t *card,
                          const int32_t *node, const int64_t *off, const int32_t *parents,
                          int64_t F, int32_t metric, int32_t nthreads, double *out) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t f = 0; f < F; ++f)
        out[f] = oracle_family_score(codes, N, stride, card, node[f], parents + off[f],
                                     (int32_t)(off[f + 1] - off[f]), metric);
    return 0;
}

/* B DAGs as dense adjacency [B, n, n] uint8, row = parent, column = child
 * (bnlearn.py:44 / bnlearn_score.R:7-13,35).  Every DAG recounts all of its n families. */
int oracle_score_dags_adj(const uint8_t *codes, int64_t N, int64_t stride, int32_t n,
                          const int32_t *card, const uint8_t *adj, int64_t B, int32_t metric,
                          int32_t nthreads, double *out) {
    double *fam = (double *)malloc((size_t)(B * n) * sizeof(double));
    if (!fam) return -1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t t = 0; t < B * n; ++t) {
        int64_t b = t / n;
        int32_t i = (int32_t)(t % n);
        int32_t parents[1024];
        int32_t k = 0;
        const uint8_t *a = adj + b * (int64_t)n * n;
        for (int32_t p = 0; p < n && k < 1024; ++p)
            if (a[(int64_t)p * n + i]) parents[k++] = p;
        fam[t] = oracle_family_score(codes, N, stride, card, i, parents, k, metric);
    }
    for (int64_t b = 0; b < B; ++b) {
        double s = 0.0;
        for (int32_t i = 0; i < n; ++i) s += fam[b * n + i];
        out[b] = s;
    }
    free(fam);
    return 0;
}
