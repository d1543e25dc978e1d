"""CPU: the C-ABI library loads and exports what include/bicgpu.h declares; host-side logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import dags_vae_search_b200 as pkg
from dags_vae_search_b200 import _native as nat
from dags_vae_search_b200 import dist as bdist
from dags_vae_search_b200 import synth, wire
from oracle import bic_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    nat.build()
    return nat.lib()


def test_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "bicgpu.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char \*)\s*(bic_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(nat.SYMBOLS), declared ^ set(nat.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bic_version() == 200
    assert "sm_100a" in nat.build_info() and "ABI 200" in nat.build_info()


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", nat.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    ctx = ctypes.c_void_p()
    rc = lib.bic_create(ctypes.byref(ctx), 0)
    assert rc == -1 and not ctx.value
    assert b"no CPU fallback" in lib.bic_last_error(None)
    with pytest.raises(pkg.BicError):
        pkg.BNLearnWrapper("asia", "bic")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "dags_vae_search_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "liboracle" not in src, f


def test_unknown_metric_and_dataset():
    with pytest.raises(NotImplementedError):
        pkg.BNLearnWrapper("asia", "mbde")
    with pytest.raises(ValueError):
        pkg.load_dataset("no_such_dataset")
    codes, card, names = pkg.load_dataset("asia")
    assert codes.shape == (8, 5000) and list(card) == [2] * 8 and names == list("ASTLBEXD")
    codes, card, names = pkg.load_dataset("sachs")
    assert codes.shape == (11, 5000) and list(card) == [3] * 11


def test_csv_loader_matches_sorted_levels(tmp_path):
    p = tmp_path / "d.csv"
    p.write_text("A,B\nyes,LOW\nno,HIGH\nyes,AVG\nno,LOW\n")
    codes, card, names = pkg.load_csv(str(p))
    assert names == ["A", "B"] and list(card) == [2, 3]
    assert codes.tolist() == [[1, 0, 1, 0], [2, 1, 0, 2]]


def test_wire_roundtrip_against_oracle(known_answer, golden_dir):
    import pyarrow.parquet as pq
    d = known_answer["graph_dict"]
    labels, ebits = wire.pack_dicts([d], 8)
    assert np.array_equal(wire.to_adjacency(labels, ebits)[0], O.labeled_dict_to_adjacency(d, 8))
    t = pq.read_table(os.path.join(golden_dir, "labeled_sample.parquet"))
    labels, ebits = wire.pack_table(t, 8)
    rows = t.to_pylist()
    l2, e2 = wire.pack_dicts(rows, 8)
    assert np.array_equal(labels, l2) and np.array_equal(ebits, e2)
    adj = wire.to_adjacency(labels, ebits)
    for b, row in enumerate(rows):
        assert np.array_equal(adj[b], O.labeled_dict_to_adjacency(row, 8))
    fixture = np.load(os.path.join(golden_dir, "asia_candidates_10k.npz"))
    assert np.array_equal(fixture["labels"][:64], labels) and np.array_equal(fixture["ebits"][:64], ebits[:, :, 0])


def test_wire_any_n_and_decoder_adapter_cpu():
    """Host logic of rows a8 / f2 / f3 without a GPU: multi-word edge masks and uint16 labels
    round-trip for n > 32, and the decoder adapter's torch packing equals pack_dicts."""
    import torch
    from dags_vae_search_b200 import decode_adapter
    for n in (8, 33, 70):
        dags = synth.er_candidates(n, 40, n - 1, 2 * n, 4, seed=n)
        labels, ebits = wire.from_adjacency(dags)
        assert labels.dtype == np.uint16 and ebits.dtype == np.uint32 and ebits.shape == (40, n, wire.edge_words(n))
        assert np.array_equal(wire.to_adjacency(labels, ebits), dags)
        dicts = [{**{f"l{i}": int(labels[b, i]) for i in range(n)},
                  **{f"e{i}": "".join(str((int(ebits[b, i, u >> 5]) >> (u & 31)) & 1) for u in range(i)) for i in range(n)}}
                 for b in range(40)]
        l2, e2 = wire.pack_dicts(dicts, n)
        assert np.array_equal(l2, labels) and np.array_equal(e2, ebits)
        # decoder-shaped tensors: PACE types (label + 3) and lower-triangular edge draws
        draws = np.zeros((40, n, n), dtype=bool)
        for b in range(40):
            for i in range(n):
                for u in range(i):
                    draws[b, i, u] = (int(ebits[b, i, u >> 5]) >> (u & 31)) & 1
        draws_noise = draws | np.triu(np.ones((n, n), dtype=bool))          # u >= v must be ignored
        lt, et = decode_adapter.decoded_to_wire(torch.from_numpy(labels.astype(np.int64) + 3), torch.from_numpy(draws_noise), n)
        assert np.array_equal(lt.numpy(), labels.astype(np.int32))
        assert np.array_equal(et.numpy().view(np.uint32), ebits)
        bad = labels.astype(np.int64) + 3
        bad[0, 0] = 1                                                        # output node as a vertex type
        lt, _ = decode_adapter.decoded_to_wire(torch.from_numpy(bad), torch.from_numpy(draws), n)
        assert lt[0, 0] == 65535
    # labels outside uint16 never wrap into range
    l3, _ = wire.pack_dicts([{"l0": 70000, "l1": 1, "e0": "", "e1": "0"}], 2)
    assert l3[0, 0] == 65535
    st = decode_adapter.DecodeState(5, 4, "cpu")
    for idx in range(2, 6):
        t = st.step(idx, torch.zeros(5, 7), torch.ones(5, idx - 1, 1), force_type=idx + 1)
        assert (t == idx + 1).all()
    lab, eb = st.wire()
    assert lab.tolist() == [[0, 1, 2, 3]] * 5 and eb[:, :, 0].tolist() == [[0, 1, 3, 7]] * 5    # every earlier vertex is a parent


def test_synth_candidates_are_dags():
    adj = synth.er_candidates(12, 200, 11, 26, 4, seed=3)
    assert adj.shape == (200, 12, 12) and adj.sum(axis=1).max() <= 4
    assert all(O.is_acyclic(a) for a in adj)
    net, card, cpts = synth.make_network(12, 20, 4, [2], seed=42)
    assert net.sum() == 20 and O.is_acyclic(net)
    codes = synth.forward_sample(net, card, cpts, 5000, np.random.default_rng(1))
    assert codes.shape == (12, 5000) and codes.max() == 1
    moves = synth.local_moves(net, 20, 3, 4, seed=5)
    assert all(O.is_acyclic(a) for a in moves) and moves.sum(axis=1).max() <= 4


def test_shard_range_partitions():
    for total in (0, 1, 7, 100, 4097):
        for world in (1, 2, 3, 8):
            spans = [bdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r"""
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from dags_vae_search_b200 import dist as bdist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
total = 11
lo, hi = bdist.shard_range(total, rank, 2)
full = bdist.gather_scores(np.arange(lo, hi, dtype=np.float64) * -1.5, total)
assert np.array_equal(full, np.arange(total) * -1.5), full
uid = bdist.broadcast_unique_id()
assert len(uid) == 128
import torch
t = torch.tensor(list(uid), dtype=torch.int64)
dist.all_reduce(t)
assert (t == 2 * torch.tensor(list(uid), dtype=torch.int64)).all()   # both ranks hold the same id
dist.destroy_process_group()
print("ok")
"""


def test_candidate_sharding_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0 and "ok" in o, o


def test_predictor_dataset_builder(tmp_path, golden_dir):
    """Row f1: the caller of the scorer (reference src/predictors/utils.py:15-59), batched.  Fake
    encoder + fake evaluator: no GPU needed to check batching, parquet schema and partitioning."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    from dags_vae_search_b200 import predictors

    class Model:
        def __init__(self):
            self.calls = 0

        def encode(self, graphs):
            self.calls += 1
            return np.array([[float(g), 2.0 * g] for g in graphs], dtype=np.float32), None

    class Evaluator:
        def __init__(self):
            self.batch_calls = 0

        def score(self, g):
            return -1.5 * g

        def score_batch(self, graphs):
            self.batch_calls += 1
            return np.array([self.score(g) for g in graphs])

    model, ev = Model(), Evaluator()
    loader = [[0, 1, 2], [3, 4, 5], [6]]
    out = str(tmp_path / "predictor_dataset")
    n = predictors.create_predictor_dataset(model, loader, out, ev.score, npartitions=2)   # bound method, as main.py:295
    assert n == 7 and model.calls == 3 and ev.batch_calls == 3
    files = sorted(os.listdir(out))
    assert files == ["part.0.parquet", "part.1.parquet"] and not os.path.exists(out + "_tmp")
    t = pa.concat_tables([pq.read_table(os.path.join(out, f)) for f in files])
    assert t.schema.field("vector").type == pa.list_(pa.float32()) and t.schema.field("target").type == pa.float64()
    assert t.column("target").to_pylist() == [-1.5 * g for g in range(7)]
    assert t.column("vector").to_pylist()[3] == [3.0, 6.0]
    # a plain callable evaluator (no score_batch) and a one-graph-at-a-time encoder still work
    class OneAtATime:
        def encode(self, graphs):
            assert len(graphs) == 1
            return np.array([[float(graphs[0])]], dtype=np.float32), None
    rows = predictors.generate_predictor_graphs_batch(OneAtATime(), lambda g: float(g), [4, 5])
    assert [r["target"] for r in rows] == [4.0, 5.0] and rows[1]["vector"].tolist() == [5.0]


def test_bench_workload_definitions():
    """Every bench workload builds its candidates and its config text on the CPU (a typo here only
    shows up on the GPU box otherwise)."""
    import bench
    assert bench.host_threads() >= 1
    for name, cfg in bench.WORKLOADS.items():
        assert isinstance(bench.describe_candidates(cfg), str)
        small = 6 if not cfg.get("row_sharded") else 3
        adj = bench.candidate_batch(cfg, small, 1, 0, 2)
        assert adj.shape == (small, cfg["n"], cfg["n"]) and adj.dtype == np.uint8
        assert all(O.is_acyclic(a) for a in adj[:3])
    cfg = bench.WORKLOADS["asia"]
    _, card, codes = bench.make_dataset_cpu(cfg, 20_000)
    assert codes.shape == (8, 20_000) and codes.max() <= 1
    # the re-sampled asia rows keep the marginals of the bundled data within sampling noise
    base = pkg.load_dataset("asia")[0]
    assert np.abs(codes.mean(axis=1) - base.mean(axis=1)).max() < 0.02


# ---------------------------------------------------------------- launch planning (no GPU)
def test_class3_range_plan():
    """How a table above one CTA's shared memory is cut into sub-ranges (range_plan in
    csrc/common.cuh, shared by the kernels and exported as bic_range_plan): by cell index in steps
    of 49 152 cells, or along the states of the first parent when the table below that parent has
    at most 16 383 cells AND that needs no more passes; 16-bit counters double the sub-range."""
    rp = nat.range_plan
    # 21^4 cells (diabetes-shaped in-degree 3): 4 sub-ranges by cell index; along the first parent it would be 5 passes of <= 5 states
    assert rp(194481, 3, 21) == {"span": 49152, "passes": 4, "states_per_pass": 0}
    assert rp(194481, 3, 21, counters16=True) == {"span": 98304, "passes": 2, "states_per_pass": 0}
    # 250 x 240: two passes either way -> runs of 125 states
    assert rp(60000, 1, 250) == {"span": 30000, "passes": 2, "states_per_pass": 125}
    assert rp(105000, 2, 250) == {"span": 35280, "passes": 3, "states_per_pass": 84}
    assert rp(50400, 4, 21) == {"span": 26400, "passes": 2, "states_per_pass": 11}
    assert rp(201600, 5, 21) == {"span": 48000, "passes": 5, "states_per_pass": 5}
    # 100 800 cells below a 3-state first parent: too large for 16-bit lanes, cut by cell index
    assert rp(302400, 6, 3) == {"span": 49152, "passes": 7, "states_per_pass": 0}
    # more than six parents: the generic row loop, 32-bit counters even in the 16-bit variant
    assert rp(177147, 10, 3) == {"span": 49152, "passes": 4, "states_per_pass": 0}
    assert rp(177147, 10, 3, counters16=True) == {"span": 49152, "passes": 4, "states_per_pass": 0}
    # every plan covers the table exactly once
    for cells, k, rad0 in [(194481, 3, 21), (60000, 1, 250), (279300, 4, 20), (55860, 3, 21), (4000000, 3, 250)]:
        for c16 in (False, True):
            p = rp(cells, k, rad0, counters16=c16)
            assert (p["passes"] - 1) * p["span"] < cells <= p["passes"] * p["span"]
            if p["states_per_pass"]:
                assert p["span"] == p["states_per_pass"] * (cells // rad0) and p["span"] <= 49152
    with pytest.raises(nat.BicError):
        rp(0, 1, 2)


def test_launch_plan_rules(monkeypatch):
    """bic_plan_slices is the host arithmetic the library uses to cut a batch of new families into
    count-kernel work items (csrc/bicgpu.cu:plan_count).  It never changes a result, only the time,
    so what is pinned here are its rules, on the shapes of the BASELINE configs."""
    for v in ("BIC_SLICE_MODEL", "BIC_RANGE_PASSES", "BIC_L2_WINDOW_MB", "BIC_L2_WINDOW_MAX_MB", "BIC_CLUSTER", "BIC_CLUSTER_SIZE"):
        monkeypatch.delenv(v, raising=False)
    MB = 1 << 20

    def windows(n, N, mb):
        return -(-n * N // (mb * MB))

    # alarm-shaped step: 17.5 k small-table families run one after another on 370 MB of rows ->
    # 32 MB L2 windows; no class-3 table
    N = 10_000_000
    alarm = nat.plan_slices(N, 37, [(5, 700)] * 17500 + [(6, 4000)] * 1800 + [(6, 16384)])
    assert alarm["slices"][0] == windows(37, N, 32) and not alarm["ranged"] and alarm["passes"] == 1
    # the same step when every column streams from the 2-bit packed copy: 4x the rows per window
    packed = nat.plan_slices(N, 37, [(5, 700)] * 17500 + [(6, 4000)] * 1800 + [(6, 16384)], all_packed=True)
    assert packed["slices"][0] == windows(37, N, 128)
    assert all(1 <= s <= N // 65536 for s in alarm["slices"])
    # a handful of families of a search step: enough slices to occupy the GPU, well below the window count of a full pass
    # (tables averaging >= 256 cells run 512-thread class-0 CTAs, two per SM; small tables 256-thread CTAs, four per SM)
    few = nat.plan_slices(N, 37, [(5, 700)] * 40)
    assert 148 * 2 // 40 <= few["slices"][0] <= 4 * 148 * 2 // 40
    few_small = nat.plan_slices(N, 37, [(3, 81)] * 40)
    assert 148 * 4 // 40 <= few_small["slices"][0] <= 4 * 148 * 4 // 40

    # diabetes-shaped local moves (5.2 GB of rows): the few large-table families are not cut into
    # 162 L2 windows (their merges would cost more than the counting); class 3 (194 481 cells) runs
    # in 4 sub-range passes, or with BIC_CLUSTER=1 in one pass over clusters of 4 CTAs
    N = 12_500_000
    fams = [(2, 600)] * 240 + [(2, 6000)] * 99 + [(3, 30000)] * 33 + [(3, 194481)] * 7
    monkeypatch.setenv("BIC_CLUSTER", "1")
    diab = nat.plan_slices(N, 413, fams, tables_in_hbm=True)
    assert diab["cluster"] == 4 and not diab["ranged"] and diab["slices"][3] <= N // (4 * 194481)
    monkeypatch.setenv("BIC_CLUSTER_SIZE", "8")
    assert nat.plan_slices(N, 413, fams, tables_in_hbm=True)["cluster"] == 8
    monkeypatch.delenv("BIC_CLUSTER_SIZE")
    assert nat.plan_slices(N, 20, [(8, 49152 * 8)])["cluster"] == 8 and nat.plan_slices(N, 20, [(3, 49153)])["cluster"] == 2
    assert nat.plan_slices(N, 20, [(8, 49152 * 8 + 1)])["cluster"] == 0      # beyond 8 x 192 KB of distributed shared memory
    monkeypatch.delenv("BIC_CLUSTER")
    diab = nat.plan_slices(N, 413, fams, tables_in_hbm=True)
    assert diab["ranged"] and diab["passes"] == 4 and diab["cluster"] == 0
    assert diab["slices"][1] < 20 and diab["slices"][2] < 10 and diab["slices"][3] <= N // (4 * 194481)
    assert windows(413, N, 256) <= diab["slices"][0] < windows(413, N, 32)      # all resident at once: wide windows
    pigs = nat.plan_slices(N, 441, [(2, 81)] * 391, tables_in_hbm=True)
    assert windows(441, N, 256) <= pigs["slices"][0] <= windows(441, N, 128)
    # the same families with the model off: the 32 MB rule of kernel versions a-h
    monkeypatch.setenv("BIC_SLICE_MODEL", "0")
    old = nat.plan_slices(N, 413, fams, tables_in_hbm=True)
    assert old["slices"][:3] == [windows(413, N, 32)] * 3
    monkeypatch.delenv("BIC_SLICE_MODEL")

    # few rows: never sliced, and a table much larger than the row count goes straight to HBM
    sachs = nat.plan_slices(5000, 11, [(4, 243)] * 7000 + [(7, 6561)] * 2000 + [(10, 177147)])
    assert sachs["slices"] == [1, 1, 1, 1] and not sachs["ranged"]
    # sub-range passes only up to BIC_RANGE_PASSES sub-ranges and with >= 4 rows per cell
    assert nat.plan_slices(N, 20, [(8, 49152 * 8)])["passes"] == 8
    assert not nat.plan_slices(N, 20, [(8, 49152 * 8 + 1)])["ranged"]
    assert not nat.plan_slices(4 * 194481 - 1, 20, [(3, 194481)])["ranged"]
    monkeypatch.setenv("BIC_RANGE_PASSES", "0")
    assert not nat.plan_slices(N, 413, fams)["ranged"]
    monkeypatch.delenv("BIC_RANGE_PASSES")

    # argument checks
    bad = nat.PlanIn(sm_count=0, N=10, n=2)
    out = nat.PlanOut()
    assert nat.lib().bic_plan_slices(ctypes.byref(bad), ctypes.byref(out)) == -2
