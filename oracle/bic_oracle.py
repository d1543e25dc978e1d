"""CPU oracle for the BIC score path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``dags_vae_search_b200``) never imports it and has no CPU fallback.

What it restates
----------------
The reference score path is ``BNLearnWrapper.score`` (reference
``src/problem/bn/bnlearn.py:27-61``) which spawns
``src/problem/bn/bnlearn_scripts/bnlearn_score.R`` whose line 38 calls
``bnlearn::score(net, dataset, type="bic")``.  The arithmetic lives in the CRAN R
package **bnlearn** — a third-party dependency that is neither vendored under
``/root/reference`` nor version-pinned anywhere in it (``requirements*.txt`` pin
Python packages only) and R itself is absent from this image.  This module restates
bnlearn's published decomposable discrete BIC:

    BIC(G) = sum_i [ sum_{j,k : N_ijk>0} N_ijk * ln(N_ijk / N_ij)
                     - 0.5 * ln(N) * (r_i - 1) * q_i ]

with natural log, ``q_i`` the product of the *declared* cardinalities of the parents
of node ``i`` (unobserved parent configurations still pay the penalty) and zero cells
skipped.

Parity status
-------------
* **asia / bic: PINNED** against the reference's own artefacts: the known-answer test
  ``tests/problem/bn/test_bnlearn.py:55`` (-13331.093616667435) and the 1408 BIC
  values the reference wrote to ``experiments/01_bn_asia/predictor_dataset/part-*.parquet``
  (``src/predictors/utils.py:24-31``).  ``tests/test_oracle_golden.py`` replays both from
  ``tests/golden/`` (made by ``tools/make_golden.py``).
* **sachs, synthetic_v12_c2, alarm-, diabetes-, pigs-shaped data and every metric
  other than "bic": parity unpinned** — the reference holds no test, fixture or
  output for them (bnlearn ships no ``sachs`` data set, so the reference scorer could
  not even load it, ``bnlearn_score.R:25-26``).

Conventions fixed here (the reference is silent; needed for bit-exact count tables)
-----------------------------------------------------------------------------------
* variable ``i`` = i-th dataset column (``bnlearn_score.R:29`` builds the graph from
  ``names(dataset)``); adjacency row = parent, column = child (``bnlearn.py:44``,
  ``bnlearn_score.R:35``).
* state code = rank of the level string in sorted order (R factor / ``np.unique``).
* parents sorted ascending by variable index, first parent most significant:
  ``j = (..(x_p1 * r_p2 + x_p2) * r_p3 + ..)``; table cell = ``j * r_i + x_i``.
* counts are int64 here (the CUDA path uses int32 while N < 2**31).
"""
from __future__ import annotations

import csv
import math
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np

__all__ = [
    "load_csv_codes", "family_q", "family_counts", "family_loglik", "family_score",
    "score_parent_sets", "score_adjacency", "parents_from_adjacency", "is_acyclic",
    "labeled_dict_to_adjacency",
]


# --------------------------------------------------------------------------- data
def load_csv_codes(path: str) -> Tuple[np.ndarray, np.ndarray, List[str], List[List[str]]]:
    """CSV of level strings -> (codes uint8 [n, N] column-major, card int32 [n], names, levels).

    Follows ``bnlearn_score.R:25-29``: column order defines the variable index; level
    codes follow R's factor order (sorted level strings).
    """
    with open(path, newline="") as fh:
        rows = list(csv.reader(fh))
    names = rows[0]
    body = rows[1:]
    # data/bn_asia/README.md:8-14 writes the CSV with write.csv(); the shipped file has no
    # row-name column, but tolerate one.
    if body and len(body[0]) == len(names) + 1:
        body = [r[1:] for r in body]
    n = len(names)
    N = len(body)
    codes = np.zeros((n, N), dtype=np.uint8)
    card = np.zeros(n, dtype=np.int32)
    levels: List[List[str]] = []
    for c in range(n):
        col = np.array([r[c] for r in body])
        lv, inv = np.unique(col, return_inverse=True)
        if len(lv) > 255:
            raise ValueError(f"column {names[c]} has {len(lv)} levels (> 255)")
        codes[c] = inv.astype(np.uint8)
        card[c] = len(lv)
        levels.append([str(x) for x in lv])
    return codes, card, names, levels


# ------------------------------------------------------------------------ families
def _norm_parents(parents: Iterable[int]) -> List[int]:
    return sorted(int(p) for p in parents)


def family_q(card: np.ndarray, parents: Iterable[int]) -> int:
    """q_i = product of declared parent cardinalities (python int, unbounded)."""
    q = 1
    for p in _norm_parents(parents):
        q *= int(card[p])
    return q


def family_config_index(codes: np.ndarray, card: np.ndarray, parents: Iterable[int]) -> np.ndarray:
    """Mixed-radix parent configuration index per row, int64 [N]."""
    N = codes.shape[1]
    j = np.zeros(N, dtype=np.int64)
    for p in _norm_parents(parents):
        j = j * int(card[p]) + codes[p].astype(np.int64)
    return j


def family_counts(codes: np.ndarray, card: np.ndarray, node: int, parents: Iterable[int]) -> np.ndarray:
    """Dense contingency table N_ijk as int64 [q, r]; cell (j, k) = j * r + k."""
    r = int(card[node])
    q = family_q(card, parents)
    cell = family_config_index(codes, card, parents) * r + codes[node].astype(np.int64)
    return np.bincount(cell, minlength=q * r).astype(np.int64).reshape(q, r)


def family_loglik(counts: np.ndarray) -> float:
    """sum_{jk: N_ijk>0} N_ijk * ln(N_ijk / N_ij) in fp64."""
    nij = counts.sum(axis=1, keepdims=True)
    mask = counts > 0
    if not mask.any():
        return 0.0
    nijk = counts[mask].astype(np.float64)
    den = np.broadcast_to(nij, counts.shape)[mask].astype(np.float64)
    return float(np.sum(nijk * np.log(nijk / den)))


def metric_penalty(metric: str, N: int, r: int, q: int) -> float:
    """Penalty subtracted from the family log-likelihood, per bnlearn ``score(type=...)``."""
    nparams = float((r - 1)) * float(q)
    if metric == "bic":
        return 0.5 * math.log(N) * nparams if N > 0 else 0.0
    if metric == "aic":
        return nparams
    if metric == "loglik":
        return 0.0
    raise NotImplementedError(f"metric {metric!r}")


def family_bd(counts: np.ndarray, a_ijk: float) -> float:
    """Bayesian-Dirichlet family term (bnlearn "bde" = BDeu with a_ijk = iss / (q r), "k2" with
    a_ijk = 1): sum_j [lgamma(a_ij) - lgamma(a_ij + N_ij) + sum_k (lgamma(a_ijk + N_ijk) - lgamma(a_ijk))]."""
    from scipy.special import gammaln
    r = counts.shape[1]
    a_ij = a_ijk * r
    nij = counts.sum(axis=1).astype(np.float64)
    c = counts.astype(np.float64)
    return float(np.sum(gammaln(a_ij) - gammaln(a_ij + nij)) + np.sum(gammaln(a_ijk + c) - gammaln(a_ijk)))


def family_score(codes: np.ndarray, card: np.ndarray, node: int, parents: Iterable[int],
                 metric: str = "bic", iss: float = 1.0) -> float:
    N = codes.shape[1]
    counts = family_counts(codes, card, node, parents)
    if metric in ("bde", "k2"):
        q, r = counts.shape
        return family_bd(counts, iss / (q * r) if metric == "bde" else 1.0)
    return family_loglik(counts) - metric_penalty(metric, N, int(card[node]), family_q(card, parents))


# ---------------------------------------------------------------------------- DAGs
def parents_from_adjacency(adj: np.ndarray) -> List[List[int]]:
    """adj [n, n], row = parent, col = child (``bnlearn.py:44``) -> parent list per node."""
    adj = np.asarray(adj)
    return [list(np.flatnonzero(adj[:, i])) for i in range(adj.shape[1])]


def is_acyclic(adj: np.ndarray) -> bool:
    """Kahn peel; ``amat(net) <- adj`` (``bnlearn_score.R:35``) rejects cyclic input."""
    adj = (np.asarray(adj) != 0)
    alive = np.ones(adj.shape[0], dtype=bool)
    while alive.any():
        indeg = adj[alive][:, :].sum(axis=0)
        removable = alive & (indeg == 0)
        if not removable.any():
            return False
        alive &= ~removable
        adj = adj & alive[:, None]
    return True


def score_parent_sets(codes: np.ndarray, card: np.ndarray, parent_sets: Sequence[Iterable[int]],
                      metric: str = "bic", cache: Dict | None = None) -> float:
    """Decomposable score of one DAG given its parent set per node (sum in node order)."""
    total = 0.0
    for i, ps in enumerate(parent_sets):
        key = (i, tuple(_norm_parents(ps)))
        if cache is not None and key in cache:
            s = cache[key]
        else:
            s = family_score(codes, card, i, key[1], metric)
            if cache is not None:
                cache[key] = s
        total += s
    return total


def score_adjacency(codes: np.ndarray, card: np.ndarray, adj: np.ndarray, metric: str = "bic",
                    cache: Dict | None = None) -> float:
    return score_parent_sets(codes, card, parents_from_adjacency(adj), metric, cache)


# ------------------------------------------------------------- candidate wire format
def labeled_dict_to_adjacency(d: Dict, n: int) -> np.ndarray:
    """``l*/e*`` dict (reference ``src/toolkit/labeled.py:132-154``) -> adjacency in
    BN-variable space as ``BNLearnWrapper.score`` builds it (``bnlearn.py:38-44``):
    vertex ``i`` carries label ``l_i``; ``e_i[u] == 1`` means edge ``u -> i``; the scorer
    relabels to ``adj[l_u, l_i] = 1`` (row = parent)."""
    adj = np.zeros((n, n), dtype=np.uint8)
    for i in range(n):
        e = d[f"e{i}"]
        li = int(d[f"l{i}"])
        for u in range(i):
            if int(e[u]) == 1:
                adj[int(d[f"l{u}"]), li] = 1
    return adj
