"""Candidate wire format of the reference (``src/toolkit/labeled.py:116-185``) without igraph.

On disk / in ``LabeledDag`` dicts a DAG over ``n`` vertices in topological order is ``l0..l{n-1}``
(vertex label = BN variable, uint16) and ``e0..e{n-1}`` (``e_i`` = string or list of ``i`` 0/1
flags, ``e_i[u] == 1`` <=> edge vertex u -> vertex i).  The scorer ingests it as two small
integer arrays: ``labels[B, n]`` (uint16) and ``ebits[B, n, EW]`` (uint32, ``EW = ceil(n / 32)``
words per vertex) with bit ``u % 32`` of word ``u // 32`` of ``ebits[b, i]`` = ``e_i[u]``; the
relabel of ``bnlearn.py:38-42`` happens on the GPU (``bic_score_dags_wire16``).  Any n <= 1024.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import numpy as np


def edge_words(n: int) -> int:
    return (int(n) + 31) // 32


def _pack_flags(flags: np.ndarray, n: int) -> np.ndarray:
    """flags uint8/bool [B, i] (i <= n) -> uint32 [B, EW]."""
    B, i = flags.shape
    EW = edge_words(n)
    padded = np.zeros((B, EW * 32), dtype=np.uint32)
    padded[:, :i] = flags
    weights = (np.uint32(1) << np.arange(32, dtype=np.uint32))
    return (padded.reshape(B, EW, 32) * weights).sum(axis=2, dtype=np.uint32)


def pack_dicts(dicts: Iterable[Dict], n: int) -> Tuple[np.ndarray, np.ndarray]:
    """``LabeledDag`` dicts -> (labels uint16 [B, n], ebits uint32 [B, n, EW])."""
    dicts = list(dicts)
    EW = edge_words(n)
    labels = np.zeros((len(dicts), n), dtype=np.uint16)
    ebits = np.zeros((len(dicts), n, EW), dtype=np.uint32)
    for b, d in enumerate(dicts):
        for i in range(n):
            lab = int(d[f"l{i}"])
            labels[b, i] = lab if 0 <= lab <= 65535 else 65535     # out of range stays out of range (rejected)
            e = d[f"e{i}"]
            if len(e) != i:
                raise ValueError(f"{i} elements expected to be in 'e{i}'")   # labeled.py:109-112
            for u in range(i):
                if int(e[u]) == 1:
                    ebits[b, i, u >> 5] |= np.uint32(1 << (u & 31))
    return labels, ebits


def pack_table(table, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """pyarrow Table in the reference parquet schema (``labeled.py:116-130``) -> packed arrays,
    column-wise (no per-row Python objects)."""
    B = table.num_rows
    EW = edge_words(n)
    labels = np.zeros((B, n), dtype=np.uint16)
    ebits = np.zeros((B, n, EW), dtype=np.uint32)
    for i in range(n):
        col = np.asarray(table.column(f"l{i}").to_numpy())
        labels[:, i] = np.where((col < 0) | (col > 65535), 65535, col)
        if i == 0 or B == 0:
            continue
        col = table.column(f"e{i}").combine_chunks()
        raw = np.frombuffer(col.buffers()[2], dtype=np.uint8)   # string data: B*i chars
        offs = np.frombuffer(col.buffers()[1], dtype=np.int32, count=B + 1, offset=col.offset * 4)
        chars = raw[offs[0]:offs[-1]]
        if chars.size != B * i:
            raise ValueError(f"{i} elements expected to be in 'e{i}'")
        ebits[:, i, :] = _pack_flags(chars.reshape(B, i) - ord("0"), n)
    return labels, ebits


def _words(ebits: np.ndarray, B: int, n: int) -> np.ndarray:
    ebits = np.asarray(ebits, dtype=np.uint32)
    return ebits.reshape(B, n, -1)


def to_adjacency(labels: np.ndarray, ebits: np.ndarray) -> np.ndarray:
    """Packed wire arrays -> adjacency uint8 [B, n, n] in BN-variable space (row = parent), the
    matrix ``BNLearnWrapper.score`` serialises (``bnlearn.py:38-44``).  Host-side convenience;
    the scorer's wire entry point does not need it."""
    labels = np.asarray(labels).astype(np.int64)
    B, n = labels.shape
    ew = _words(ebits, B, n)
    adj = np.zeros((B, n, n), dtype=np.uint8)
    rows = np.arange(B)
    for i in range(n):
        for u in range(i):
            m = ((ew[:, i, u >> 5] >> np.uint32(u & 31)) & np.uint32(1)).astype(bool)
            adj[rows[m], labels[m, u], labels[m, i]] = 1
    return adj


def from_adjacency(adj: np.ndarray, rng=None) -> Tuple[np.ndarray, np.ndarray]:
    """Adjacency uint8 [B, n, n] of DAGs (row = parent) -> wire arrays: vertices in a topological
    order of each DAG (``LabeledDag.from_graph_to_dict``, ``labeled.py:156-172``), labels = the BN
    variables in that order.  Host-side; used by tests and the bench to feed wide networks through
    the wire entry point."""
    adj = np.asarray(adj, dtype=np.uint8)
    B, n, _ = adj.shape
    EW = edge_words(n)
    labels = np.zeros((B, n), dtype=np.uint16)
    ebits = np.zeros((B, n, EW), dtype=np.uint32)
    for b in range(B):
        a = adj[b].astype(bool)
        indeg = a.sum(axis=0)
        order, ready = [], [v for v in range(n) if indeg[v] == 0]
        while ready:
            v = ready.pop(0)
            order.append(v)
            for c in np.flatnonzero(a[v]):
                indeg[c] -= 1
                if indeg[c] == 0:
                    ready.append(int(c))
        if len(order) != n:
            raise ValueError("graph has a cycle")
        pos = np.empty(n, dtype=np.int64)
        pos[order] = np.arange(n)
        labels[b] = order
        for i, v in enumerate(order):
            for p in np.flatnonzero(a[:, v]):
                u = int(pos[p])
                ebits[b, i, u >> 5] |= np.uint32(1 << (u & 31))
    return labels, ebits


def dict_to_edges(d: Dict, n: int) -> List[Tuple[int, int]]:
    """Edges (u, i) in vertex space of one ``LabeledDag`` dict (``labeled.py:132-154``)."""
    return [(u, i) for i in range(n) for u in range(i) if int(d[f"e{i}"][u]) == 1]
