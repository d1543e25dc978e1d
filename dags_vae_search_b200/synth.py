"""Synthetic networks, datasets and candidate DAGs for the BASELINE configs that ship no data
(SURVEY.md section 8d): random DAG of a named shape -> Dirichlet CPTs -> forward-sampled rows,
and Erdos-Renyi candidate DAGs following the reference's recipe
(``src/toolkit/labeled.py:281-333``: G(n, m) undirected, oriented low -> high vertex, random
label permutation).  numpy only; seeds make every config reproducible.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def random_dag(n: int, e: int, max_indegree: int, rng: np.random.Generator) -> np.ndarray:
    """Adjacency uint8 [n, n] (row = parent) with exactly ``e`` edges where possible, acyclic by
    a random vertex order, in-degree capped."""
    order = rng.permutation(n)
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    pairs = [(u, v) for u in range(n) for v in range(n) if pos[u] < pos[v]]
    rng.shuffle(pairs)
    adj = np.zeros((n, n), dtype=np.uint8)
    indeg = np.zeros(n, dtype=np.int64)
    left = e
    for u, v in pairs:
        if left == 0:
            break
        if indeg[v] < max_indegree:
            adj[u, v] = 1
            indeg[v] += 1
            left -= 1
    return adj


def random_cpts(adj: np.ndarray, card: np.ndarray, rng: np.random.Generator, alpha: float = 1.0) -> List[np.ndarray]:
    """One table [q_i, r_i] per node, rows ~ Dirichlet(alpha)."""
    n = adj.shape[0]
    cpts = []
    for i in range(n):
        ps = np.flatnonzero(adj[:, i])
        q = int(np.prod(card[ps])) if len(ps) else 1
        cpts.append(rng.dirichlet(np.full(int(card[i]), alpha), size=q))
    return cpts


def topo_order(adj: np.ndarray) -> List[int]:
    adj = adj.astype(bool).copy()
    n = adj.shape[0]
    alive = np.ones(n, dtype=bool)
    order: List[int] = []
    while alive.any():
        ready = np.flatnonzero(alive & ~adj[alive].any(axis=0))
        if len(ready) == 0:
            raise ValueError("graph has a cycle")
        order.extend(int(x) for x in ready)
        alive[ready] = False
    return order


def forward_sample(adj: np.ndarray, card: np.ndarray, cpts: List[np.ndarray], N: int,
                   rng: np.random.Generator, chunk: int = 1 << 20) -> np.ndarray:
    """uint8 codes [n, N]; parent configuration index = mixed radix over ascending parents, first
    most significant (the scorer's convention)."""
    n = adj.shape[0]
    codes = np.zeros((n, N), dtype=np.uint8)
    order = topo_order(adj)
    for s in range(0, N, chunk):
        m = min(chunk, N - s)
        for i in order:
            ps = np.flatnonzero(adj[:, i])
            j = np.zeros(m, dtype=np.int64)
            for p in ps:
                j = j * int(card[p]) + codes[p, s:s + m]
            cum = np.cumsum(cpts[i], axis=1)
            u = rng.random(m)
            x = (u[:, None] > cum[j, :-1]).sum(axis=1)
            codes[i, s:s + m] = x.astype(np.uint8)
    return codes


def make_network(n: int, e: int, max_indegree: int, card_choices, seed: int, alpha: float = 1.0):
    rng = np.random.default_rng(seed)
    adj = random_dag(n, e, max_indegree, rng)
    card = rng.choice(np.asarray(card_choices), size=n).astype(np.int32)
    cpts = random_cpts(adj, card, rng, alpha)
    return adj, card, cpts


def er_candidates(n: int, B: int, m_lo: int, m_hi: int, max_indegree: Optional[int], seed: int) -> np.ndarray:
    """B Erdos-Renyi candidate DAGs as adjacency uint8 [B, n, n] in BN-variable space.

    Recipe of the reference generator (``labeled.py:281-333``): pick m of the n(n-1)/2 vertex
    pairs, orient each low -> high vertex index, then give the vertices a random permutation of
    the variable labels.  ``m`` is uniform in [m_lo, m_hi]; parents beyond ``max_indegree`` are
    dropped at random."""
    rng = np.random.default_rng(seed)
    iu, iv = np.triu_indices(n, k=1)          # u < v: edge u -> v
    P = len(iu)
    m = rng.integers(m_lo, m_hi + 1, size=B)
    score = rng.random((B, P))
    kth = np.sort(score, axis=1)[np.arange(B), np.minimum(m, P) - 1]
    keep = score <= kth[:, None]
    vert = np.zeros((B, n, n), dtype=np.uint8)
    vert[:, iu, iv] = keep
    if max_indegree is not None:
        w = rng.random((B, n, n)) * vert               # random priority per present edge
        thresh = -np.sort(-w, axis=1)[:, min(max_indegree, n) - 1, :] if max_indegree < n else np.zeros((B, n))
        vert = (vert.astype(bool) & (w >= np.maximum(thresh[:, None, :], 1e-300))).astype(np.uint8)
    perm = np.argsort(rng.random((B, n)), axis=1)       # vertex -> variable label
    adj = np.zeros_like(vert)
    bidx = np.arange(B)[:, None, None]
    adj[bidx, perm[:, :, None], perm[:, None, :]] = vert
    return adj


def local_moves(true_adj: np.ndarray, B: int, max_moves: int, max_indegree: int, seed: int) -> np.ndarray:
    """B neighbours of ``true_adj``: up to ``max_moves`` random edge additions / deletions /
    reversals each, kept acyclic and within the in-degree cap (a local-search candidate set)."""
    rng = np.random.default_rng(seed)
    n = true_adj.shape[0]
    out = np.zeros((B, n, n), dtype=np.uint8)
    for b in range(B):
        a = true_adj.copy()
        for _ in range(int(rng.integers(1, max_moves + 1))):
            u, v = rng.choice(n, size=2, replace=False)
            c = a.copy()
            if c[u, v]:
                c[u, v] = 0
                if rng.random() < 0.5:
                    c[v, u] = 1
            elif not c[v, u]:
                c[u, v] = 1
            if c.sum(axis=0).max() > max_indegree:
                continue
            try:
                topo_order(c)
            except ValueError:
                continue
            a = c
        out[b] = a
    return out
