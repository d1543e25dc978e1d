"""dags_vae_search_b200 — B200-native BIC structure scorer.

Drop-in for the score path of rlog58/dags-vae-search (``src/problem/bn``): same
``BNLearnWrapper(dataset, metric).score(graph) -> float``, backed by hand-written sm_100a
kernels behind a C ABI (``include/bicgpu.h``).  No CPU fallback.
"""
from ._native import BicError, build
from .bnlearn import LABEL_KEY, BNLearnWrapper
from .datasets import load_csv, load_dataset, register_dataset
from .scorer import BicScorer

__all__ = ["BNLearnWrapper", "BicScorer", "BicError", "LABEL_KEY", "build", "load_csv", "load_dataset",
           "register_dataset"]
