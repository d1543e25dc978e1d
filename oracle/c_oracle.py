"""ctypes binding of oracle/liboracle.so (oracle/bic_oracle.c) — TEST INFRASTRUCTURE ONLY.

Used by tests (cross-check against the numpy oracle, parity checker at sizes numpy is too
slow for) and by bench.py's cpu_baseline / --impl reference legs (the timed CPU stand-in for
the reference's Rscript-per-DAG path, reference src/problem/bn/bnlearn.py:46-54, which cannot
run here: R, bnlearn, igraph and pgmpy are absent).  Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METRICS = {"bic": 0, "loglik": 1, "aic": 2, "bde": 3, "k2": 4}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "bic_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i32p = ctypes.POINTER(ctypes.c_int32)
        i64p = ctypes.POINTER(ctypes.c_int64)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.oracle_max_threads.restype = ctypes.c_int
        L.oracle_set_iss.argtypes = [ctypes.c_double]
        L.oracle_set_iss.restype = None
        L.oracle_family_counts.argtypes = [u8p, ctypes.c_int64, ctypes.c_int64, i32p, ctypes.c_int32,
                                           i32p, ctypes.c_int32, i64p]
        L.oracle_family_counts.restype = ctypes.c_int
        L.oracle_score_families.argtypes = [u8p, ctypes.c_int64, ctypes.c_int64, i32p, i32p, i64p, i32p,
                                            ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, f64p]
        L.oracle_score_families.restype = ctypes.c_int
        L.oracle_score_dags_adj.argtypes = [u8p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, i32p, u8p,
                                            ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, f64p]
        L.oracle_score_dags_adj.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def set_iss(iss: float) -> None:
    lib().oracle_set_iss(float(iss))


def max_threads() -> int:
    return int(lib().oracle_max_threads())


def _codes(codes):
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    return codes, codes.shape[1], codes.shape[1]


def family_counts(codes, card, node: int, parents: Sequence[int]) -> np.ndarray:
    codes, N, stride = _codes(codes)
    card = np.ascontiguousarray(card, dtype=np.int32)
    ps = np.ascontiguousarray(sorted(int(p) for p in parents), dtype=np.int32)
    q = int(np.prod([int(card[p]) for p in ps], dtype=object)) if len(ps) else 1
    r = int(card[node])
    out = np.zeros(q * r, dtype=np.int64)
    lib().oracle_family_counts(_p(codes, ctypes.c_uint8), N, stride, _p(card, ctypes.c_int32), int(node),
                               _p(ps, ctypes.c_int32), len(ps), _p(out, ctypes.c_int64))
    return out.reshape(q, r)


def score_families(codes, card, node, off, parents, metric: str = "bic", nthreads: int = 0) -> np.ndarray:
    codes, N, stride = _codes(codes)
    card = np.ascontiguousarray(card, dtype=np.int32)
    node = np.ascontiguousarray(node, dtype=np.int32)
    off = np.ascontiguousarray(off, dtype=np.int64)
    parents = np.ascontiguousarray(parents, dtype=np.int32)
    if parents.size == 0:
        parents = np.zeros(1, dtype=np.int32)
    out = np.zeros(len(node), dtype=np.float64)
    lib().oracle_score_families(_p(codes, ctypes.c_uint8), N, stride, _p(card, ctypes.c_int32),
                                _p(node, ctypes.c_int32), _p(off, ctypes.c_int64), _p(parents, ctypes.c_int32),
                                len(node), METRICS[metric], int(nthreads), _p(out, ctypes.c_double))
    return out


def score_dags_adj(codes, card, adj, metric: str = "bic", nthreads: int = 0) -> np.ndarray:
    """adj uint8 [B, n, n], row = parent; reference semantics: no family cache."""
    codes, N, stride = _codes(codes)
    card = np.ascontiguousarray(card, dtype=np.int32)
    adj = np.ascontiguousarray(adj, dtype=np.uint8)
    B, n, _ = adj.shape
    out = np.zeros(B, dtype=np.float64)
    rc = lib().oracle_score_dags_adj(_p(codes, ctypes.c_uint8), N, stride, n, _p(card, ctypes.c_int32),
                                     _p(adj, ctypes.c_uint8), B, METRICS[metric], int(nthreads),
                                     _p(out, ctypes.c_double))
    if rc != 0:
        raise MemoryError("oracle_score_dags_adj")
    return out
