#!/usr/bin/env python
"""Generate tests/golden/* from the reference tree (run in the build container only).

Reads reference DATA artefacts (never source code) under /root/reference and writes small,
compressed fixtures so that tests and bench.py never touch /root/reference at run time:

  asia.npz / sachs.npz          sorted-level uint8 codes [n, N], cardinalities, names, levels
                                (reference data/bn_asia/target.csv, data/bn_sachs/target.csv)
  asia_known_answer.json        the l*/e* dict and expected BIC of reference
                                tests/problem/bn/test_bnlearn.py:22-39,55
  asia_predictor_targets.npy    the 1408 BIC values the reference scorer wrote to
                                experiments/01_bn_asia/predictor_dataset/part-{0..21}.parquet
  asia_test_dags.npz            the 22 022 DAGs those targets were drawn from
                                (experiments/01_bn_asia/data/test/part.0.parquet), as labels + e-bits
  asia_candidates_10k.npz       first 10 000 DAGs of data/bn_asia/encoder_dataset (config 1)
  sachs_candidates_100k.npz     first 100 000 DAGs of data/bn_sachs/encoder_dataset (config 2)
  labeled_sample.parquet        64 rows of the l*/e* wire format, re-written by pyarrow here

DAG fixtures keep the reference wire format (labels l_i + edge bits e_i packed into one
integer per vertex: bit u of ebits[i] = e_i[u]) so the ingest path under test is the real one.
"""
import json
import os
import sys

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
from oracle.bic_oracle import load_csv_codes  # noqa: E402


def pack_wire(table: pa.Table, n: int, limit=None):
    if limit is not None:
        table = table.slice(0, limit)
    B = table.num_rows
    labels = np.zeros((B, n), dtype=np.uint16)
    ebits = np.zeros((B, n), dtype=np.uint32)
    for i in range(n):
        labels[:, i] = table.column(f"l{i}").to_numpy()
        col = table.column(f"e{i}").to_pylist()
        if i:
            arr = np.frombuffer("".join(col).encode(), dtype=np.uint8).reshape(B, i) - ord("0")
            assert arr.max() <= 1
            ebits[:, i] = (arr.astype(np.uint32) << np.arange(i, dtype=np.uint32)).sum(axis=1)
    return labels, ebits


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, rel in (("asia", "data/bn_asia/target.csv"), ("sachs", "data/bn_sachs/target.csv")):
        codes, card, names, levels = load_csv_codes(os.path.join(REF, rel))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), codes=codes, card=card,
                            names=np.array(names), levels=np.array(json.dumps(levels)))
        print(name, codes.shape, card)

    known = {
        "source": "reference tests/problem/bn/test_bnlearn.py:22-39,55",
        "dataset": "asia", "metric": "bic", "expected": -13331.093616667435, "abs_tol": 1e-5,
        "graph_dict": {"l0": 0, "l1": 1, "l2": 2, "l3": 3, "l4": 4, "l5": 5, "l6": 6, "l7": 7,
                       "e0": [], "e1": [1], "e2": [0, 0], "e3": [0, 0, 0], "e4": [0, 1, 0, 0],
                       "e5": [1, 1, 0, 0, 0], "e6": [0, 1, 0, 0, 1, 0], "e7": [0, 0, 0, 1, 1, 1, 0]},
    }
    with open(os.path.join(OUT, "asia_known_answer.json"), "w") as fh:
        json.dump(known, fh, indent=1)

    targets = []
    for i in range(22):
        t = pq.read_table(os.path.join(REF, f"experiments/01_bn_asia/predictor_dataset/part-{i}.parquet"))
        targets.append(t.column("target").to_numpy())
    targets = np.concatenate(targets).astype(np.float64)
    assert targets.shape == (1408,)
    np.save(os.path.join(OUT, "asia_predictor_targets.npy"), targets)

    t = pq.read_table(os.path.join(REF, "experiments/01_bn_asia/data/test/part.0.parquet"))
    labels, ebits = pack_wire(t, 8)
    np.savez_compressed(os.path.join(OUT, "asia_test_dags.npz"), labels=labels.astype(np.uint8),
                        ebits=ebits.astype(np.uint8))
    print("asia test dags", labels.shape)

    t = pq.read_table(os.path.join(REF, "data/bn_asia/encoder_dataset/part.0.parquet"))
    labels, ebits = pack_wire(t, 8, 10000)
    np.savez_compressed(os.path.join(OUT, "asia_candidates_10k.npz"), labels=labels.astype(np.uint8),
                        ebits=ebits.astype(np.uint8))
    # a few rows in the original on-disk format for the wire-format ingest test
    cols = [f"l{i}" for i in range(8)] + [f"e{i}" for i in range(8)]
    pq.write_table(t.slice(0, 64).select(cols), os.path.join(OUT, "labeled_sample.parquet"))

    t = pq.read_table(os.path.join(REF, "data/bn_sachs/encoder_dataset/part.0.parquet"))
    labels, ebits = pack_wire(t, 11, 100000)
    np.savez_compressed(os.path.join(OUT, "sachs_candidates_100k.npz"), labels=labels.astype(np.uint8),
                        ebits=ebits.astype(np.uint16))
    print("sachs candidates", labels.shape)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
