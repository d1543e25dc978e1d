"""Candidate wire format of the reference (``src/toolkit/labeled.py:116-185``) without igraph.

On disk / in ``LabeledDag`` dicts a DAG over ``n`` vertices in topological order is ``l0..l{n-1}``
(vertex label = BN variable, uint16) and ``e0..e{n-1}`` (``e_i`` = string or list of ``i`` 0/1
flags, ``e_i[u] == 1`` <=> edge vertex u -> vertex i).  The scorer ingests it as two small
integer arrays: ``labels[B, n]`` and ``ebits[B, n]`` with bit u of ``ebits[b, i]`` = ``e_i[u]``;
the relabel of ``bnlearn.py:38-42`` happens on the GPU (``bic_score_dags_wire``).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import numpy as np


def pack_dicts(dicts: Iterable[Dict], n: int) -> Tuple[np.ndarray, np.ndarray]:
    """``LabeledDag`` dicts -> (labels uint8 [B, n], ebits uint32 [B, n]).  n <= 32."""
    dicts = list(dicts)
    labels = np.zeros((len(dicts), n), dtype=np.uint8)
    ebits = np.zeros((len(dicts), n), dtype=np.uint32)
    for b, d in enumerate(dicts):
        for i in range(n):
            labels[b, i] = int(d[f"l{i}"])
            e = d[f"e{i}"]
            if len(e) != i:
                raise ValueError(f"{i} elements expected to be in 'e{i}'")   # labeled.py:109-112
            word = 0
            for u in range(i):
                if int(e[u]) == 1:
                    word |= 1 << u
            ebits[b, i] = word
    return labels, ebits


def pack_table(table, n: int) -> Tuple[np.ndarray, np.ndarray]:
    """pyarrow Table in the reference parquet schema (``labeled.py:116-130``) -> packed arrays,
    column-wise (no per-row Python objects)."""
    B = table.num_rows
    labels = np.zeros((B, n), dtype=np.uint8)
    ebits = np.zeros((B, n), dtype=np.uint32)
    for i in range(n):
        labels[:, i] = table.column(f"l{i}").to_numpy()
        if i == 0 or B == 0:
            continue
        col = table.column(f"e{i}").combine_chunks()
        raw = np.frombuffer(col.buffers()[2], dtype=np.uint8)   # string data: B*i chars
        offs = np.frombuffer(col.buffers()[1], dtype=np.int32, count=B + 1, offset=col.offset * 4)
        chars = raw[offs[0]:offs[-1]]
        if chars.size != B * i:
            raise ValueError(f"{i} elements expected to be in 'e{i}'")
        flags = (chars.reshape(B, i) - ord("0")).astype(np.uint32)
        ebits[:, i] = (flags << np.arange(i, dtype=np.uint32)).sum(axis=1, dtype=np.uint32)
    return labels, ebits


def to_adjacency(labels: np.ndarray, ebits: np.ndarray) -> np.ndarray:
    """Packed wire arrays -> adjacency uint8 [B, n, n] in BN-variable space (row = parent), the
    matrix ``BNLearnWrapper.score`` serialises (``bnlearn.py:38-44``).  Host-side convenience;
    the scorer's wire entry point does not need it."""
    labels = np.asarray(labels)
    ebits = np.asarray(ebits, dtype=np.uint32)
    B, n = labels.shape
    adj = np.zeros((B, n, n), dtype=np.uint8)
    rows = np.arange(B)
    for i in range(n):
        for u in range(i):
            m = ((ebits[:, i] >> np.uint32(u)) & np.uint32(1)).astype(bool)
            adj[rows[m], labels[m, u], labels[m, i]] = 1
    return adj


def dict_to_edges(d: Dict, n: int) -> List[Tuple[int, int]]:
    """Edges (u, i) in vertex space of one ``LabeledDag`` dict (``labeled.py:132-154``)."""
    return [(u, i) for i in range(n) for u in range(i) if int(d[f"e{i}"][u]) == 1]
