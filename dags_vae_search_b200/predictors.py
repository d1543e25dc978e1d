"""Batched predictor-dataset builder — the production caller of the scorer.

Mirrors the reference's ``src/predictors/utils.py:15-59`` (``generate_predictor_graphs_batch`` /
``create_predictor_dataset``, called from ``experiments/01_bn_asia/main.py:268-303``): for every
graph of a loader batch, the VAE mean ``mu`` and the evaluator's score go into one parquet row
``{vector: list<float>, target: double}``.  Differences, all on the caller side of the score path:

* the evaluator is called once per *batch* (``score_batch``) when it offers that, instead of
  once per graph (the reference spawns one Rscript per graph, ``utils.py:24``);
* the encoder is called once per batch when it accepts that, else per graph as in the reference;
* the parts written to ``<output_dir>_tmp`` are finally consolidated into ``npartitions`` files
  under ``output_dir`` (the reference accepts ``npartitions`` but never uses it, ``utils.py:37-59``).

The VAE itself is out of scope: ``model`` is anything with ``encode(list_of_graphs) -> (mu, _)``.
"""
from __future__ import annotations

import logging
import os
import shutil
from typing import Iterable, List

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

logger = logging.getLogger(__name__)


def _to_numpy(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach()
    if hasattr(x, "cpu"):
        x = x.cpu()
    return np.asarray(x.numpy() if hasattr(x, "numpy") else x, dtype=np.float32)


def _encode_batch(model, graphs: List) -> np.ndarray:
    try:
        mu, _ = model.encode(graphs)
        mu = _to_numpy(mu)
        if mu.ndim == 2 and mu.shape[0] == len(graphs):
            return mu
    except Exception:   # encoder that only takes one graph at a time (reference usage, utils.py:23)
        pass
    return np.stack([_to_numpy(model.encode([g])[0])[0] for g in graphs])


def _score_batch(evaluator, graphs: List) -> np.ndarray:
    owner = getattr(evaluator, "__self__", evaluator)     # bound BNLearnWrapper.score -> the wrapper
    if hasattr(owner, "score_batch"):
        return np.asarray(owner.score_batch(graphs), dtype=np.float64)
    return np.array([evaluator(g) for g in graphs], dtype=np.float64)


def generate_predictor_graphs_batch(model, evaluator, graphs: Iterable) -> List[dict]:
    graphs = list(graphs)
    if not graphs:
        return []
    Z = _encode_batch(model, graphs)
    y = _score_batch(evaluator, graphs)
    return [{"vector": Z[i], "target": float(y[i])} for i in range(len(graphs))]


_SCHEMA = pa.schema([pa.field("vector", pa.list_(pa.float32())), pa.field("target", pa.float64())])


def _to_table(rows: List[dict]) -> pa.Table:
    return pa.table({"vector": pa.array([r["vector"].tolist() for r in rows], type=pa.list_(pa.float32())),
                     "target": pa.array([r["target"] for r in rows], type=pa.float64())}, schema=_SCHEMA)


def create_predictor_dataset(model, graphs_dataloader, output_dir: str, evaluator, npartitions: int = 4) -> int:
    """Returns the number of rows written."""
    tmp_dir = output_dir + "_tmp"
    os.makedirs(tmp_dir, exist_ok=True)
    parts = []
    for batch_ind, batch in enumerate(graphs_dataloader):
        rows = generate_predictor_graphs_batch(model, evaluator, batch)
        path = f"{tmp_dir}/part-{batch_ind}.parquet"
        pq.write_table(_to_table(rows), path)
        parts.append(path)
    tables = [pq.read_table(p) for p in parts]
    total = sum(t.num_rows for t in tables)
    os.makedirs(output_dir, exist_ok=True)
    if tables:
        full = pa.concat_tables(tables)
        npartitions = max(1, min(int(npartitions), max(total, 1)))
        bounds = np.linspace(0, total, npartitions + 1).astype(int)
        for k in range(npartitions):
            pq.write_table(full.slice(bounds[k], bounds[k + 1] - bounds[k]), f"{output_dir}/part.{k}.parquet")
    shutil.rmtree(tmp_dir, ignore_errors=True)
    logger.info("Dataset creation completed.")
    return total
