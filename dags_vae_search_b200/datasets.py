"""Discrete datasets for the scorer: named bundles, CSV loading, registration.

The reference resolves ``dataset_name`` inside R (``data(list = dataset_name)``,
``bnlearn_score.R:25-26``) and only uses pgmpy's example model for the node count
(``bnlearn.py:21,28``).  Neither R nor pgmpy exist here, so names resolve to bundled state-code
arrays: ``asia`` = the reference's ``data/bn_asia/target.csv`` (bnlearn's ``asia``, 5000 x 8),
``sachs`` = ``data/bn_sachs/target.csv`` (5000 x 11).  Codes are the rank of the level string in
sorted order (R factor order); column order defines the variable index (``bnlearn_score.R:29``).
"""
from __future__ import annotations

import csv
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DATA = os.path.join(_HERE, "data")
_REGISTRY: Dict[str, Tuple[np.ndarray, np.ndarray, List[str]]] = {}


def register_dataset(name: str, codes, card, names: Optional[List[str]] = None) -> None:
    """Make ``BNLearnWrapper(name, ...)`` resolve to caller-supplied data (uint8 ``[n, N]``)."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    card = np.ascontiguousarray(card, dtype=np.int32)
    if codes.ndim != 2 or codes.shape[0] != card.shape[0]:
        raise ValueError("codes must be [n, N] with one cardinality per variable")
    _REGISTRY[name] = (codes, card, list(names) if names is not None else [f"V{i}" for i in range(len(card))])


def load_csv(path: str) -> Tuple[np.ndarray, np.ndarray, List[str]]:
    """CSV of level strings with a header row -> (codes uint8 [n, N], card int32 [n], names)."""
    with open(path, newline="") as fh:
        rows = list(csv.reader(fh))
    names, body = rows[0], rows[1:]
    if body and len(body[0]) == len(names) + 1:   # write.csv() row-name column
        body = [r[1:] for r in body]
    table = np.array(body)
    n = len(names)
    codes = np.zeros((n, table.shape[0]), dtype=np.uint8)
    card = np.zeros(n, dtype=np.int32)
    for v in range(n):
        levels, inv = np.unique(table[:, v], return_inverse=True)
        if len(levels) > 255:
            raise ValueError(f"column {names[v]} has {len(levels)} levels (> 255)")
        codes[v] = inv
        card[v] = len(levels)
    return codes, card, names


def load_dataset(name: str) -> Tuple[np.ndarray, np.ndarray, List[str]]:
    """Registered name, bundled name (``asia``, ``sachs``), or a path to ``.npz`` / ``.csv``."""
    if name in _REGISTRY:
        return _REGISTRY[name]
    bundled = os.path.join(_DATA, f"{name}.npz")
    path = bundled if os.path.exists(bundled) else name
    if path.endswith(".npz") and os.path.exists(path):
        d = np.load(path)
        return (np.ascontiguousarray(d["codes"], dtype=np.uint8), np.ascontiguousarray(d["card"], dtype=np.int32),
                [str(x) for x in d["names"]])
    if path.endswith(".csv") and os.path.exists(path):
        return load_csv(path)
    raise ValueError(f"unknown dataset {name!r}: register_dataset() it, or pass a .npz/.csv path "
                     f"(bundled: {sorted(f[:-4] for f in os.listdir(_DATA) if f.endswith('.npz'))})")
