"""GPU: round-2 parity tests — stream ordering of device inputs, the wire format for any n, the
decoder adapter (row f3), the predictor-dataset builder on the real scorer (row f1), the fused
reduce-scatter path of row-sharded runs on one GPU, and the bench datasets at full size
(diabetes-, pigs-shaped 12.5 M rows, synthetic_v12_c2) against the C oracle.

Counts are bit-exact; scores agree within 1e-9 relative (BASELINE.json north_star)."""
import ctypes
import os

import numpy as np
import pytest

import dags_vae_search_b200 as pkg
from dags_vae_search_b200 import _native as nat
from dags_vae_search_b200 import decode_adapter, predictors, synth, wire
from oracle import bic_oracle as O
from oracle import c_oracle as C
from graph_stub import Graph

pytestmark = pytest.mark.gpu
RTOL = 1e-9   # north_star: BIC within 1e-9 relative in fp64


def assert_scores(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    rel = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-300)
    assert rel.size == 0 or rel.max() <= RTOL, rel.max()


# ------------------------------------------------------------ stream ordering (ADVICE, high)
def test_device_inputs_are_ordered_after_their_producer(asia):
    """A CUDA tensor handed to the scorer may still be being written by torch's stream: a long
    kernel is queued right before the copy that fills `adj`, then score_adjacency is called at
    once.  The library runs on its own stream and must wait for the producer (bic_wait_stream)."""
    import torch
    codes, card = asia
    dags = synth.er_candidates(8, 3000, 5, 14, None, seed=5)
    src = torch.from_numpy(dags).cuda()
    lab_h, eb_h = wire.from_adjacency(dags[:500])
    lab_src, eb_src = torch.from_numpy(lab_h.astype(np.int32)).cuda(), torch.from_numpy(eb_h.astype(np.int64)).cuda()
    with pkg.BicScorer(codes, card) as s:
        want = s.score_adjacency(dags)
        side = torch.cuda.Stream()
        for _ in range(3):
            adj = torch.zeros_like(src)                  # all-empty DAGs unless the copy below has landed
            lab, eb = torch.zeros_like(lab_src), torch.zeros_like(eb_src)
            torch.cuda.synchronize()
            with torch.cuda.stream(side):
                torch.cuda._sleep(400_000_000)           # ~0.2 s of device time ahead of the producer
                adj.copy_(src, non_blocking=True)
                lab.copy_(lab_src, non_blocking=True)
                eb.copy_(eb_src, non_blocking=True)
                got = s.score_adjacency(adj, no_cache=True)
                got_w = s.score_wire(lab, eb)
            assert np.array_equal(got.cpu().numpy(), want)
            assert np.array_equal(got_w.cpu().numpy(), want[:500])
        # a dataset uploaded from a CUDA tensor that is still being produced
        dev_codes = torch.zeros((8, codes.shape[1]), dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(side):
            torch.cuda._sleep(200_000_000)
            dev_codes.copy_(torch.from_numpy(codes).cuda(), non_blocking=True)
            with pkg.BicScorer(dev_codes, card) as s2:
                assert np.array_equal(s2.score_adjacency(dags[:64]), want[:64])


# ------------------------------------------------------- wire format for any n (rows a8 / f2)
@pytest.mark.parametrize("n,N", [(8, 3000), (12, 4000), (32, 3000), (33, 3000), (37, 20000), (100, 2000), (441, 1500)])
def test_wire_format_any_n(n, N):
    """labels uint16 + ceil(n / 32) edge words per vertex (labeled.py:116-130) through
    bic_score_dags_wire16: same bits as the adjacency entry point, host and device inputs;
    labels that are not a permutation (also >= 256: no uint8 wrap-around) are rejected."""
    import torch
    adj_t, card, cpts = synth.make_network(n, min(2 * n, n * (n - 1) // 2), 3, [2, 3, 4], seed=n)
    codes = synth.forward_sample(adj_t, card, cpts, N, np.random.default_rng(n + 1))
    dags = synth.er_candidates(n, 300, n - 1, 2 * n, 4, seed=n + 2)
    labels, ebits = wire.from_adjacency(dags)
    assert labels.dtype == np.uint16 and ebits.shape == (300, n, (n + 31) // 32)
    assert np.array_equal(wire.to_adjacency(labels, ebits), dags)
    with pkg.BicScorer(codes, card) as s:
        want = s.score_adjacency(dags, no_cache=True)
        assert_scores(want[:20], C.score_dags_adj(codes, card, dags[:20]))
        got, bad = s.score_wire(labels, ebits, no_cache=True, return_invalid=True)
        assert bad == 0 and np.array_equal(got, want)
        dev = s.score_wire(torch.from_numpy(labels.astype(np.int32)).cuda(), torch.from_numpy(ebits.astype(np.int64)).cuda())
        assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)
        # bits at u >= v are ignored (labeled.py:143-145 only reads e_i[0..i-1])
        noisy = ebits.copy()
        noisy[:, 0, :] = 0xFFFFFFFF
        for v in range(1, n):
            noisy[:, v, v >> 5] |= np.uint32((0xFFFFFFFF << (v & 31)) & 0xFFFFFFFF)
            noisy[:, v, (v >> 5) + 1:] = 0xFFFFFFFF
        assert np.array_equal(s.score_wire(labels, noisy), want)
        # invalid label sets
        badlab = labels.astype(np.int64).copy()
        badlab[0, 1] = badlab[0, 0]            # repeated variable
        badlab[1, 0] = n                       # out of range
        badlab[2, 0] = badlab[2, 0] + 256      # would wrap to a valid label in uint8
        badlab[3, 0] = 70000                   # beyond uint16
        got, bad = s.score_wire(badlab, ebits, return_invalid=True)
        assert bad == 4 and np.isnan(got[:4]).all() and np.array_equal(got[4:], want[4:])


# ----------------------------------------------------------- decoder adapter (row f3)
def reference_decode_bookkeeping(types, decisions, n):
    """The graph bookkeeping of PaceVaeV3.decode (pace.py:1684-1741) and
    from_pace_graph_to_labeled_graph (pace.py:1290-1305) restated with plain Python lists for ONE
    candidate: types[t] = sampled PACE type at step idx = t + 2, decisions[t][vi] = edge draw for
    PACE vertex vi + 1 -> idx.  Returns (labels, edges) of the labeled graph, or None where the
    reference cannot build one (early output node)."""
    OUT = 1
    pace_types = [2, 0]                       # start sign, input node
    edges = set()
    for t in range(n):
        idx = t + 2
        if types[t] == OUT:
            return None                       # finished early: from_pace_graph_to_labeled_graph indexes missing vertices
        pace_types.append(types[t])
        for vi in range(idx - 2, -1, -1):
            if decisions[t][vi]:
                edges.add((vi + 1, idx))
    labels = [pace_types[v] - 3 for v in range(2, n + 2)]
    ledges = [(u - 2, v - 2) for (u, v) in sorted(edges) if u >= 2]
    return labels, ledges


@pytest.mark.parametrize("n", [8, 12, 37, 100])
def test_decoder_adapter_matches_host_path(n):
    """Random decoder-shaped tensors -> DecodeState / decoded_to_wire -> score_wire (CUDA in, CUDA
    out) equals the host path (reference bookkeeping -> LabeledDag dict -> pack_dicts ->
    to_adjacency -> score_adjacency) bit for bit; candidates the reference cannot score are NaN."""
    import torch
    B, N = 256, 3000
    rng = np.random.default_rng(n)
    adj_t, card, cpts = synth.make_network(n, min(2 * n, n * (n - 1) // 2), 3, [2, 3], seed=n)
    codes = synth.forward_sample(adj_t, card, cpts, N, np.random.default_rng(n + 1))
    types = np.stack([rng.permutation(n) + 3 for _ in range(B)])            # PACE types of the n real vertices
    p_edge = min(0.5, 3.0 / n)
    decisions = rng.random((B, n, n + 1)) < p_edge                           # [b, t, vi], vi <= t (idx - 2)
    types[0, 2] = 1                        # output node sampled early
    types[1, 0] = 2                        # start sign as a vertex type
    types[2, 3] = types[2, 1]              # a variable twice
    # the adapter's input: edge_draws[b, v, u] = decisions[b, v, u + 1]
    draws = np.zeros((B, n, n), dtype=bool)
    for v in range(n):
        draws[:, v, :v] = decisions[:, v, 1:v + 1]
    with pkg.BicScorer(codes, card) as s:
        lab_t, eb_t = decode_adapter.decoded_to_wire(torch.from_numpy(types).cuda(), torch.from_numpy(draws).cuda(), n)
        assert lab_t.is_cuda and eb_t.shape == (B, n, (n + 31) // 32)
        got = s.score_wire(lab_t, eb_t, no_cache=True)
        assert got.is_cuda
        got = got.cpu().numpy()
        dicts, valid = [], []
        for b in range(B):
            ref = reference_decode_bookkeeping(list(types[b]), decisions[b], n)
            ok = ref is not None and sorted(ref[0]) == list(range(n))        # bnlearn.py:34-35
            valid.append(ok)
            if ok:
                labels, edges = ref
                eset = set(edges)
                dicts.append({**{f"l{i}": labels[i] for i in range(n)},
                              **{f"e{i}": [int((u, i) in eset) for u in range(i)] for i in range(n)}})
        valid = np.array(valid)
        assert not valid[:3].any() and valid[3:].all()
        hl, he = wire.pack_dicts(dicts, n)
        want = s.score_adjacency(wire.to_adjacency(hl, he), no_cache=True)
        assert np.isnan(got[~valid]).all()
        assert np.array_equal(got[valid], want)
        assert_scores(want[:10], C.score_dags_adj(codes, card, wire.to_adjacency(hl, he)[:10]))

        # DecodeState.step: the same draws the reference makes (categorical type, Bernoulli edges), on the device
        gen = torch.Generator(device="cuda")
        gen.manual_seed(7)
        st = decode_adapter.DecodeState(B, n, "cuda", generator=gen)
        remaining = torch.ones((B, n + 3), dtype=torch.bool, device="cuda")
        remaining[:, :3] = False
        for idx in range(2, n + 2):
            logits = torch.where(remaining, torch.zeros((), device="cuda"), torch.full((), -1e9, device="cuda"))
            es = torch.full((B, idx - 1, 1), p_edge, device="cuda")
            new = st.step(idx, logits, es)
            remaining[torch.arange(B, device="cuda"), new] = False           # a permutation, like a well-trained decoder
        lab2, eb2 = st.wire()
        sc = s.score_wire(lab2, eb2)
        assert not torch.isnan(sc).any()
        adj2 = wire.to_adjacency(lab2.cpu().numpy(), eb2.cpu().numpy().view(np.uint32))
        assert np.array_equal(sc.cpu().numpy(), s.score_adjacency(adj2))
        assert 0.2 * p_edge < adj2.mean() * n / (n - 1) * 2 < 1.8 * p_edge   # Bernoulli(p) on the n(n-1)/2 forward pairs


# ------------------------------------- predictor-dataset builder on the real scorer (row f1)
def test_predictor_dataset_reproduces_reference_targets(golden_dir, tmp_path):
    """The reference's own artefact: experiments/01_bn_asia/predictor_dataset/part-*.parquet holds
    1408 {vector, target} rows written by create_predictor_dataset(model, loader, ..,
    BNLearnWrapper("asia", "bic").score) (src/predictors/utils.py:15-59, main.py:268-303) for
    graphs of data/test.  Feeding all 22 022 test DAGs (batches of 64, stub graphs, a fake
    encoder) through the same function with the CUDA scorer as evaluator must write every one of
    the 1408 reference targets, in the reference's schema."""
    import pyarrow as pa
    import pyarrow.parquet as pq
    d = np.load(os.path.join(golden_dir, "asia_test_dags.npz"))
    labels, ebits = d["labels"], d["ebits"].astype(np.uint32)
    targets = np.load(os.path.join(golden_dir, "asia_predictor_targets.npy"))
    assert len(targets) == 1408
    n = 8
    graphs = []
    for b in range(len(labels)):
        edges = [(u, i) for i in range(n) for u in range(i) if (int(ebits[b, i]) >> u) & 1]
        graphs.append(Graph(n, edges, [int(x) for x in labels[b]]))
    loader = [graphs[i:i + 64] for i in range(0, len(graphs), 64)]

    class FakeEncoder:                      # the VAE is out of scope: mu = a function of the graph
        def encode(self, gs):
            return np.array([[float(len(g.get_edgelist())), float(g.vs["type"][0])] for g in gs], dtype=np.float32), None

    ev = pkg.BNLearnWrapper("asia", "bic")
    out = str(tmp_path / "predictor_dataset")
    rows = predictors.create_predictor_dataset(FakeEncoder(), loader, out, ev.score, npartitions=22)
    assert rows == len(graphs)
    files = sorted(os.listdir(out))
    assert len(files) == 22 and not os.path.exists(out + "_tmp")
    t = pa.concat_tables([pq.read_table(os.path.join(out, f)) for f in files])
    assert t.schema.names == ["vector", "target"]                                   # utils.py:24-31
    assert t.schema.field("vector").type == pa.list_(pa.float32()) and t.schema.field("target").type == pa.float64()
    written = np.sort(np.asarray(t.column("target").to_numpy()))
    pos = np.clip(np.searchsorted(written, targets), 1, len(written) - 1)
    nearest = np.where(np.abs(written[pos] - targets) < np.abs(written[pos - 1] - targets), written[pos], written[pos - 1])
    assert (np.abs(nearest - targets) / np.abs(targets)).max() < RTOL
    # graph by graph: the written target is the oracle's BIC of that graph
    adj = wire.to_adjacency(labels[:300], ebits[:300])
    codes, card, _ = pkg.load_dataset("asia")
    assert_scores(np.asarray(t.column("target").to_numpy())[:300], C.score_dags_adj(codes, card, adj))
    assert t.column("vector").to_pylist()[5] == [float(len(graphs[5].get_edgelist())), float(labels[5][0])]


# ------------------------- fused reduce-scatter of row-sharded count tables, on one GPU
@pytest.mark.parametrize("derive_rows", [False, True])
def test_exchange_buffer_path_world1(monkeypatch, derive_rows):
    """BIC_PUSH_WORLD1=1: a one-rank communicator takes the row-sharded exchange-buffer path
    (count kernels store the finished tables into the owner's slot, barrier all-reduce, owner sums
    the slots inside the fp64 reduce, terms travel with one all-reduce) — every kernel of the
    multi-GPU path except the IPC mapping.  Same bits as the fused single-GPU path; with derived
    families the owner's summed tables are written back for k_derive."""
    n, N = (8, 1_100_000) if derive_rows else (14, 250_000)
    adj, card, cpts = synth.make_network(n, 12 if derive_rows else 20, 3, [2, 3, 4] if not derive_rows else [2, 3], seed=8)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(9))
    dags = synth.er_candidates(n, 400, n - 1, 2 * n, 5, seed=10)
    big = np.array([21, 20, 19, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3], dtype=np.int32)
    with pkg.BicScorer(codes, card) as s:
        fused = s.score_adjacency(dags)
    monkeypatch.setenv("BIC_PUSH_WORLD1", "1")
    with pkg.BicScorer(codes, card) as s:
        buf = (ctypes.c_uint8 * 128)()
        assert nat.lib().bic_comm_unique_id(ctypes.addressof(buf)) == 0
        s.init_row_sharding(0, 1, bytes(buf))
        s.profile_enable(True)
        s.profile_reset()
        got = s.score_adjacency(dags)
        prof = s.profile()
        assert np.array_equal(got, fused)
        assert prof["exchange_ms"] > 0 and prof["exchange_fused"] > 0 and prof["exchange_nccl"] == 0
        if derive_rows:
            assert prof["families_derived"] > 0
        again = s.score_adjacency(np.concatenate([dags[:50], synth.er_candidates(n, 100, n - 1, 2 * n, 5, seed=11)]))
        assert np.array_equal(again[:50], fused[:50])
        tabs = s.count_families([3, 5], [[0, 1], [2, 4, 6, 7]])          # the caller wants tables: NCCL path
        assert np.array_equal(tabs[0], C.family_counts(codes, card, 3, [0, 1]))
        assert np.array_equal(tabs[1], C.family_counts(codes, card, 5, [2, 4, 6, 7]))
    if not derive_rows:   # class 2 / class 3 tables and several slices through the same path
        rng = np.random.default_rng(3)
        codes2 = np.stack([rng.integers(0, c, size=900_001) for c in big]).astype(np.uint8)
        fams = [(3, [0, 1]), (4, [0, 1, 2]), (0, []), (5, [6, 7])]          # 1260 (class 0), 23 940 (class 2), 21, 27 cells
        fams3 = [(3, [0, 1, 4])]                                            # 3 780 cells: class 1
        fams_big = [(4, [0, 1, 2, 5])]                                      # 71 820 cells: class 3
        with pkg.BicScorer(codes2, big) as s:
            want = s.score_families([f[0] for f in fams + fams3 + fams_big], [f[1] for f in fams + fams3 + fams_big])
            buf = (ctypes.c_uint8 * 128)()
            assert nat.lib().bic_comm_unique_id(ctypes.addressof(buf)) == 0
            s.init_row_sharding(0, 1, bytes(buf))
            got = s.score_families([f[0] for f in fams + fams3 + fams_big], [f[1] for f in fams + fams3 + fams_big])
            assert np.array_equal(got, want)
            monkeypatch.setenv("BIC_XCHG_MB", "1")                         # 262 144 cells per slot: still fits
        with pkg.BicScorer(codes2, big) as s:
            buf = (ctypes.c_uint8 * 128)()
            assert nat.lib().bic_comm_unique_id(ctypes.addressof(buf)) == 0
            s.init_row_sharding(0, 1, bytes(buf))
            many = [(4, [0, 1, 2, 5]), (4, [0, 1, 2, 6]), (4, [0, 1, 2, 7]), (4, [0, 1, 2, 8])]   # 287 280 cells > one slot: falls back
            got = s.score_families([f[0] for f in many], [f[1] for f in many])
            node = np.array([f[0] for f in many], dtype=np.int32)
            off = np.arange(0, 4 * len(many) + 1, 4, dtype=np.int64)
            par = np.array([p for f in many for p in f[1]], dtype=np.int32)
            assert_scores(got, C.score_families(codes2, big, node, off, par))


# ------------------------------------------------ the bench datasets at full size
def _bench_dataset(name, device="cuda:0"):
    import torch
    import bench
    cfg = bench.WORKLOADS[name]
    true_adj, card, codes = bench.make_dataset_gpu(cfg, cfg["rows"], torch.device(device))
    return cfg, true_adj, np.asarray(card, dtype=np.int32), codes


def _pick_families(card, targets, rng):
    """One family per requested table-size class: greedy search over random parent sets."""
    bounds = [(1, 2048), (2049, 12288), (12289, 49152), (49153, 400_000)]
    n = len(card)
    fams = []
    for cls in targets:
        lo, hi = bounds[cls]
        for _ in range(20000):
            k = int(rng.integers(1, 11))
            vs = rng.choice(n, size=k + 1, replace=False)
            cells = int(np.prod(card[vs].astype(np.int64)))
            if lo <= cells <= hi:
                fams.append((int(vs[0]), sorted(int(x) for x in vs[1:])))
                break
        else:
            raise AssertionError(f"no family of class {cls}")
    return fams


def _sub_oracle_counts(codes_dev, card, node, parents):
    """Oracle counts of one family from only the columns it touches (the full 5 GB dataset need not
    leave the GPU for a counts check)."""
    cols = [node] + list(parents)
    sub = codes_dev[cols].cpu().numpy()
    return C.family_counts(sub, card[cols], 0, list(range(1, len(cols))))


@pytest.mark.parametrize("name,classes", [("diabetes", [0, 1, 2, 3]), ("pigs", [0, 1, 2, 3])])
def test_bench_dataset_full_size_counts_and_scores(name, classes):
    """The row-sharded bench datasets exactly as bench.py builds them (413 / 441 variables x 12.5 M
    rows per GPU): counts of one family per count-kernel class against the C oracle (bit-exact),
    and the scores of the true DAG + 2 local-search neighbours (CSR entry point) within 1e-9."""
    import bench
    cfg, true_adj, card, codes = _bench_dataset(name)
    n, N = cfg["n"], cfg["rows"]
    assert tuple(codes.shape) == (n, N)
    fams = _pick_families(card, classes, np.random.default_rng(17))
    with pkg.BicScorer(codes, card) as s:
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
        for (i, ps), t in zip(fams, tabs):
            assert int(t.sum()) == N
            assert np.array_equal(t, _sub_oracle_counts(codes, card, i, ps)), (name, i, ps)
        dags = bench.candidate_batch(cfg, 3, 0, 0, 1)
        got = s.score_adjacency(dags, no_cache=True)
        host = codes.cpu().numpy()
        del codes
        assert_scores(got, C.score_dags_adj(host, card, dags))
        # the same through the wire format (n > 32: multi-word edge masks, uint16 labels)
        lab, eb = wire.from_adjacency(dags)
        assert np.array_equal(s.score_wire(lab, eb), got)


def test_bench_synthetic_v12_c2_batch():
    """BASELINE configs[2] as bench.py runs it: 100 k rows x 4096 ER candidates; every DAG score
    against the C oracle, and the unique families' counts sum to N."""
    import bench
    cfg, true_adj, card, codes = _bench_dataset("synthetic_v12_c2")
    dags = bench.candidate_batch(cfg, cfg["batch"], 3, 0, 1)
    host = codes.cpu().numpy()
    with pkg.BicScorer(codes, card) as s:
        got = s.score_adjacency(dags, no_cache=True)
        assert_scores(got, C.score_dags_adj(host, card, dags))
        st = s.cache_stats()
        assert 1000 < st["families"] <= 12 * 2 ** 11
        fams = [(int(i), [int(p) for p in np.flatnonzero(dags[b][:, i])]) for b in range(40) for i in (0, 5, 11)]
        for (i, ps), t in zip(fams, s.count_families([f[0] for f in fams], [f[1] for f in fams])):
            assert np.array_equal(t, O.family_counts(host, card, i, ps))


# ------------------------------------------- TMA-staged row tiles (experiment knob BIC_TMA=1)
@pytest.mark.parametrize("knob", ["BIC_TMA=1", "BIC_U8_NARROW=1", "BIC_U8_TWO=0", "BIC_U8_TWO=3"])
@pytest.mark.parametrize("N", [1, 15, 2047, 2048, 2049, 70_001, 600_000])
def test_tma_staged_tiles_match_default_path(N, knob, monkeypatch):
    """uint8 path of classes 0 / 1 with the rows staged through a shared-memory ring by bulk copies
    (cp.async.bulk + mbarrier, BIC_TMA=1) or loaded 8 bytes at a time (BIC_U8_NARROW=1): identical
    counts and score bits; ragged tails, 1..7 columns, all three lane modes, a family the variants
    do not serve (class 2) in the same call."""
    rng = np.random.default_rng(N)
    card = np.array([2, 3, 4, 5, 3, 2, 7, 6, 1, 3, 9, 11], dtype=np.int32)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    fams = [(0, []), (1, [0]), (2, [0, 1]), (3, [0, 1, 2]), (4, [0, 1, 2, 3]), (5, [0, 1, 2, 3, 4]),
            (9, [0, 1, 2, 3, 4, 5]), (6, [1, 3, 7]), (7, [2, 3]), (3, [8, 9]), (2, [3, 4, 6, 9]),
            (10, [6, 7, 11]), (11, [3, 6, 7, 10]), (0, [1, 2, 4, 5, 8, 9, 3])]
    node = np.array([f[0] for f in fams], dtype=np.int32)
    off = np.zeros(len(fams) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f[1]) for f in fams])
    par = np.array([p for f in fams for p in f[1]], dtype=np.int32)
    with pkg.BicScorer(codes, card) as s:
        want = s.score_families_csr(node, off, par, no_cache=True)
    monkeypatch.setenv(*knob.split("="))
    with pkg.BicScorer(codes, card) as s:
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
        for (i, ps), t in zip(fams, tabs):
            assert np.array_equal(t, O.family_counts(codes, card, i, ps)), (N, i, ps)
        assert np.array_equal(s.score_families_csr(node, off, par, no_cache=True), want)


# --------------------------------------------- cache checkpoints: content tag, key validation
def test_cache_checkpoint_refuses_other_content_and_bad_keys(sachs, tmp_path):
    """ADVICE (low): a checkpoint is tied to the dataset's CONTENT (device-side fingerprint), not only
    its shape, and to iss for bde terms; bic_cache_import validates parent masks and duplicates."""
    codes, card = sachs
    dags = synth.er_candidates(11, 200, 8, 25, 5, seed=2)
    path = str(tmp_path / "cache.npz")
    with pkg.BicScorer(codes, card) as s:
        want = s.score_adjacency(dags)
        assert s.save_cache(path) > 0
    other = codes.copy()
    other[3, 17] = (other[3, 17] + 1) % 3                     # same shape and cardinalities, one state differs
    with pkg.BicScorer(other, card) as s:
        with pytest.raises(ValueError):
            s.load_cache(path)
    with pkg.BicScorer(np.ascontiguousarray(codes[:, ::-1]), card) as s:      # the same rows in another order: another dataset
        with pytest.raises(ValueError):
            s.load_cache(path)
    with pkg.BicScorer(codes, card) as s:
        assert s.load_cache(path) > 0
        s.profile_reset()
        assert np.array_equal(s.score_adjacency(dags), want)
        assert s.profile()["families_counted"] == 0           # everything came from the checkpoint
    # bde terms depend on iss
    with pkg.BicScorer(codes, card, metric="bde", iss=2.0) as s:
        s.score_adjacency(dags[:20])
        s.save_cache(path)
    with pkg.BicScorer(codes, card, metric="bde", iss=3.0) as s:
        with pytest.raises(ValueError):
            s.load_cache(path)
    # raw import: bad parent bit, self parent, duplicate key
    with pkg.BicScorer(codes, card) as s:
        lib = nat.lib()
        terms, nparams = np.zeros(2), np.zeros(2)
        for keys in ([[0, 1 << 11], [1, 0]],        # parent index 11 >= n
                     [[2, 1 << 2], [1, 0]],         # node 2 as its own parent
                     [[11, 0], [1, 0]],             # node 11 >= n
                     [[3, 0b110000], [3, 0b110000]]):   # the same family twice
            k = np.array(keys, dtype=np.uint64)
            rc = lib.bic_cache_import(s._ctx, k.ctypes.data, terms.ctypes.data, nparams.ctypes.data, 2, 0)
            assert rc == -2, keys
            assert s.cache_stats()["families"] == 0
        k = np.array([[3, 0b110000], [4, 0b1]], dtype=np.uint64)
        assert lib.bic_cache_import(s._ctx, k.ctypes.data, terms.ctypes.data, nparams.ctypes.data, 2, 0) == 0
        assert s.cache_stats()["families"] == 2


# ----------------------------------------------------------------- edges of the new entry points
def test_wire16_at_the_variable_limit_and_flag_errors(asia):
    """n = 1024 (NMAX): 32 edge words per vertex, 17-word keys; BIC_FLAG_LOCAL_BATCH without family
    sharding and too few edge words are argument errors; bic_wait_stream accepts the legacy default
    stream (NULL)."""
    import torch
    n, N = 1024, 600
    rng = np.random.default_rng(5)
    card = rng.integers(2, 4, size=n).astype(np.int32)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    adj = np.zeros((3, n, n), dtype=np.uint8)
    for b in range(3):
        order = rng.permutation(n)
        for _ in range(1500):
            i, j = sorted(rng.choice(n, size=2, replace=False))
            if adj[b][:, order[j]].sum() < 3:
                adj[b, order[i], order[j]] = 1           # edge from an earlier to a later vertex of `order`: acyclic
    labels, ebits = wire.from_adjacency(adj)
    assert ebits.shape == (3, n, 32)
    with pkg.BicScorer(codes, card) as s:
        want = s.score_adjacency(adj)
        assert_scores(want, C.score_dags_adj(codes, card, adj))
        assert np.array_equal(s.score_wire(labels, ebits), want)
        dev = s.score_wire(torch.from_numpy(labels.astype(np.int32)).cuda(), torch.from_numpy(ebits.astype(np.int64)).cuda())
        assert np.array_equal(dev.cpu().numpy(), want)
        lib = nat.lib()
        out = np.zeros(3)
        lab16 = np.ascontiguousarray(labels, dtype=np.uint16)
        rc = lib.bic_score_dags_wire16(s._ctx, lab16.ctypes.data, ebits.ctypes.data, 31, 3, 0, out.ctypes.data, None, 0)
        assert rc == -2 and b"ewords" in lib.bic_last_error(s._ctx)
        rc = lib.bic_score_dags_adj(s._ctx, adj.ctypes.data, 3, 0, out.ctypes.data, None, nat.FLAG_LOCAL_BATCH)
        assert rc == -2 and b"family sharding" in lib.bic_last_error(s._ctx)
        assert lib.bic_wait_stream(s._ctx, None) == 0
        assert np.array_equal(s.score_adjacency(adj), want)
    codes_a, card_a = asia
    with pkg.BicScorer(codes_a, card_a) as s:      # the legacy uint8 entry point still serves n <= 32
        lab8 = np.tile(np.arange(8, dtype=np.uint8), (2, 1))
        eb = np.zeros((2, 8), dtype=np.uint32)
        eb[1, 3] = 0b101
        out = np.zeros(2)
        assert nat.lib().bic_score_dags_wire(s._ctx, lab8.ctypes.data, eb.ctypes.data, 2, 0, out.ctypes.data, None, 0) == 0
        assert np.array_equal(out, s.score_wire(lab8, eb))
