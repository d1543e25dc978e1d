"""GPU: the row-sharded (NCCL count-table all-reduce) and candidate-sharded paths."""
import os
import subprocess
import sys

import numpy as np
import pytest

import dags_vae_search_b200 as pkg
from dags_vae_search_b200 import _native as nat
from dags_vae_search_b200 import synth
from oracle import c_oracle as C

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_sharded_path_world1():
    """world = 1 communicator: exercises count-only kernels -> ncclAllReduce(uint32) -> separate
    fp64 reduce kernel on one GPU; must reproduce the fused path bit for bit."""
    import ctypes
    n, N = 14, 250_000
    adj, card, cpts = synth.make_network(n, 20, 3, [2, 3, 4], seed=8)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(9))
    dags = synth.er_candidates(n, 400, 13, 30, 6, seed=10)
    with pkg.BicScorer(codes, card) as s:
        fused = s.score_adjacency(dags)
        buf = (ctypes.c_uint8 * 128)()
        assert nat.lib().bic_comm_unique_id(ctypes.addressof(buf)) == 0
        s.init_row_sharding(0, 1, bytes(buf))
        assert s.cache_stats()["families"] == 0            # cache dropped when sharding changes
        sharded = s.score_adjacency(dags)
        assert np.array_equal(sharded, fused)
        tabs = s.count_families([3, 5], [[0, 1], [2, 4, 6, 7]])
        assert np.array_equal(tabs[0], C.family_counts(codes, card, 3, [0, 1]))
        assert np.array_equal(tabs[1], C.family_counts(codes, card, 5, [2, 4, 6, 7]))
        s.end_row_sharding()
        assert np.array_equal(s.score_adjacency(dags), fused)


def test_row_and_candidate_sharding_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    port = str(29700 + os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tests", "_multigpu_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "multigpu ok 2" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_row_sharded_with_derived_families_world1():
    import ctypes
    n, N = 8, 1_100_000
    adj, card, cpts = synth.make_network(n, 10, 3, [2, 3], seed=31)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(32))
    dags = synth.er_candidates(n, 300, 7, 16, 5, seed=33)
    with pkg.BicScorer(codes, card) as s:
        s.derive = False
        plain = s.score_adjacency(dags, no_cache=True)
        s.derive = True
        buf = (ctypes.c_uint8 * 128)()
        assert nat.lib().bic_comm_unique_id(ctypes.addressof(buf)) == 0
        s.init_row_sharding(0, 1, bytes(buf))
        s.profile_reset()
        sharded = s.score_adjacency(dags, no_cache=True)
        assert s.profile()["families_derived"] > 0
        assert np.array_equal(sharded, plain)
