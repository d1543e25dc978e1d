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


def test_two_contexts_from_two_threads(asia, sachs):
    """Contexts are independent; calls on different contexts may overlap (ctypes drops the GIL)."""
    import threading
    a_adj = synth.er_candidates(8, 4000, 5, 14, None, seed=1)
    s_adj = synth.er_candidates(11, 4000, 8, 25, None, seed=2)
    with pkg.BicScorer(*asia) as sa, pkg.BicScorer(*sachs) as ss:
        want_a = sa.score_adjacency(a_adj, no_cache=True)
        want_s = ss.score_adjacency(s_adj, no_cache=True)
        got = {}

        def work(name, scorer, adj):
            for _ in range(5):
                got[name] = scorer.score_adjacency(adj, no_cache=True)

        threads = [threading.Thread(target=work, args=("a", sa, a_adj)), threading.Thread(target=work, args=("s", ss, s_adj))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert np.array_equal(got["a"], want_a) and np.array_equal(got["s"], want_s)


def test_two_devices_in_one_process(sachs):
    """The >48 KB shared-memory opt-in of the count kernels is per device."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    codes, card = sachs
    fams_nodes = [0, 1, 2]
    fams_parents = [[1, 2, 3, 4, 5, 6, 7, 8], [0, 2, 3, 4, 5, 6, 7], [3, 4]]      # 19683 / 6561 / 27 cells
    with pkg.BicScorer(codes, card, device=0) as s0, pkg.BicScorer(codes, card, device=1) as s1:
        a = s0.score_families(fams_nodes, fams_parents)
        b = s1.score_families(fams_nodes, fams_parents)
        assert np.array_equal(a, b)
        for i, ps, v in zip(fams_nodes, fams_parents, a):
            assert v == pytest.approx(C.score_families(codes, card, *_csr1(i, ps))[0], rel=1e-9)


def _csr1(i, ps):
    return np.array([i], dtype=np.int32), np.array([0, len(ps)], dtype=np.int64), np.array(ps, dtype=np.int32)
