"""GPU: the CUDA path, called through the C ABI, against the oracle and the reference's golden
values.  Counts are bit-exact; scores agree within 1e-9 relative (BASELINE.json north_star)."""
import itertools
import math
import os

import numpy as np
import pytest

import dags_vae_search_b200 as pkg
from dags_vae_search_b200 import synth, wire
from oracle import bic_oracle as O
from oracle import c_oracle as C
from graph_stub import Graph, from_dict_to_graph

pytestmark = pytest.mark.gpu
RTOL = 1e-9   # north_star: BIC within 1e-9 relative in fp64


def assert_scores(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    rel = np.abs(got[ok] - want[ok]) / np.maximum(np.abs(want[ok]), 1e-300)
    assert rel.size == 0 or rel.max() <= RTOL, rel.max()


def csr_of(families):
    node = np.array([f[0] for f in families], dtype=np.int32)
    off = np.zeros(len(families) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(f[1]) for f in families])
    par = np.array([p for f in families for p in f[1]], dtype=np.int32)
    return node, off, par


def all_families(n, max_k=None):
    fams = []
    for i in range(n):
        others = [p for p in range(n) if p != i]
        for k in range(0, (n - 1 if max_k is None else max_k) + 1):
            for ps in itertools.combinations(others, k):
                fams.append((i, list(ps)))
    return fams


@pytest.fixture(scope="module")
def asia_scorer(asia):
    with pkg.BicScorer(*asia) as s:
        yield s


@pytest.fixture(scope="module")
def sachs_scorer(sachs):
    with pkg.BicScorer(*sachs) as s:
        yield s


# ------------------------------------------------------------------ reference golden values
def test_known_answer_through_wrapper(known_answer):
    # reference tests/problem/bn/test_bnlearn.py:46-55, igraph replaced by a stub of the same API
    graph = from_dict_to_graph(known_answer["graph_dict"], 8)
    evaluator = pkg.BNLearnWrapper("asia", "bic")
    result = evaluator.score(graph)
    assert isinstance(result, float)
    assert -13331.093616667435 == pytest.approx(result, abs=1e-5)
    assert abs(result - known_answer["expected"]) / abs(known_answer["expected"]) < RTOL
    assert evaluator(graph) == result                      # used as a bare callable (predictors/utils.py:24)
    assert evaluator.score_batch([graph, graph]).tolist() == [result, result]


def test_wrapper_error_conventions():
    ev = pkg.BNLearnWrapper("asia", "bic")
    with pytest.raises(AssertionError, match="Expected 8 vertices"):
        ev.score(Graph(7, [], list(range(7))))
    with pytest.raises(AssertionError, match="Expected graph labels from 0 to 7"):
        ev.score(Graph(8, [], [0, 1, 2, 3, 4, 5, 6, 6]))
    with pytest.raises(Exception, match="R script failed"):   # cyclic: R's amat<- refuses
        ev.score(Graph(8, [(0, 1), (1, 2), (2, 0)], list(range(8))))
    # labels permute vertices into variables (bnlearn.py:38-42)
    g1 = Graph(8, [(0, 1)], [3, 5, 0, 1, 2, 4, 6, 7])
    g2 = Graph(8, [(3, 5)], list(range(8)))
    assert ev.score(g1) == ev.score(g2)


def test_1408_reference_values(asia_scorer, golden_dir):
    targets = np.load(os.path.join(golden_dir, "asia_predictor_targets.npy"))
    d = np.load(os.path.join(golden_dir, "asia_test_dags.npz"))
    scores = asia_scorer.score_wire(d["labels"], d["ebits"].astype(np.uint32))
    assert scores.shape == (22022,) and not np.isnan(scores).any()
    order = np.sort(scores)
    pos = np.clip(np.searchsorted(order, targets), 1, len(order) - 1)
    nearest = np.where(np.abs(order[pos] - targets) < np.abs(order[pos - 1] - targets), order[pos], order[pos - 1])
    assert (np.abs(nearest - targets) / np.abs(targets)).max() < RTOL


# ----------------------------------------------------------------------- counts, bit-exact
def test_asia_all_1024_families_counts_and_scores(asia, asia_scorer):
    codes, card = asia
    fams = all_families(8)
    assert len(fams) == 1024
    tabs = asia_scorer.count_families([f[0] for f in fams], [f[1] for f in fams])
    for (i, ps), t in zip(fams, tabs):
        want = O.family_counts(codes, card, i, ps)
        assert t.dtype == np.int32 and np.array_equal(t, want), (i, ps)
    node, off, par = csr_of(fams)
    for metric in ("bic", "loglik", "aic"):
        got = asia_scorer.score_families_csr(node, off, par, metric=metric, no_cache=True)
        want = C.score_families(codes, card, node, off, par, metric=metric)
        assert_scores(got, want)


def test_sachs_families_every_class(sachs, sachs_scorer):
    """k = 0..10 parents: tables from 3 cells to 3^11 = 177 147 cells, i.e. every count-kernel
    class including the HBM-atomics one (q*r*4 B > shared memory)."""
    codes, card = sachs
    rng = np.random.default_rng(1)
    fams = []
    for k in range(0, 11):
        for _ in range(6 if k < 10 else 3):
            i = int(rng.integers(11))
            ps = sorted(rng.choice([p for p in range(11) if p != i], size=k, replace=False).tolist())
            fams.append((i, ps))
    tabs = sachs_scorer.count_families([f[0] for f in fams], [f[1] for f in fams])
    for (i, ps), t in zip(fams, tabs):
        assert np.array_equal(t, C.family_counts(codes, card, i, ps)), (i, ps)
        assert t.sum() == 5000
    node, off, par = csr_of(fams)
    assert_scores(sachs_scorer.score_families_csr(node, off, par, no_cache=True),
                  C.score_families(codes, card, node, off, par))


# --------------------------------------------------------------------------- DAG batches
def test_asia_10k_candidates(asia, asia_scorer, golden_dir):
    codes, card = asia
    d = np.load(os.path.join(golden_dir, "asia_candidates_10k.npz"))
    adj = wire.to_adjacency(d["labels"], d["ebits"].astype(np.uint32))
    true = np.zeros((1, 8, 8), dtype=np.uint8)
    for u, v in [(0, 2), (1, 3), (1, 4), (2, 5), (3, 5), (5, 6), (5, 7), (4, 7)]:
        true[0, u, v] = 1
    adj = np.concatenate([true, adj])
    want = C.score_dags_adj(codes, card, adj)
    asia_scorer.cache_clear()
    got = asia_scorer.score_adjacency(adj)
    assert_scores(got, want)
    assert got[0] == pytest.approx(-11109.741872493603, rel=RTOL)
    st = asia_scorer.cache_stats()
    assert st["lookups"] == adj.shape[0] * 8 and st["misses"] == st["families"] <= 1024
    # warm cache, cache off, and the wire-format entry point give the same bits
    assert np.array_equal(asia_scorer.score_adjacency(adj), got)
    assert np.array_equal(asia_scorer.score_adjacency(adj, no_cache=True), got)
    assert np.array_equal(asia_scorer.score_wire(d["labels"], d["ebits"].astype(np.uint32)), got[1:])


def test_sachs_100k_candidates(sachs, sachs_scorer, golden_dir):
    codes, card = sachs
    d = np.load(os.path.join(golden_dir, "sachs_candidates_100k.npz"))
    labels, ebits = d["labels"], d["ebits"].astype(np.uint32)
    got = sachs_scorer.score_wire(labels, ebits, no_cache=True)
    assert got.shape == (100000,) and not np.isnan(got).any()
    # oracle on the unique families only (the per-DAG C oracle would recount 1.1 M families)
    adj = wire.to_adjacency(labels, ebits)
    masks = (adj.astype(np.int64) << np.arange(11)[None, :, None]).sum(axis=1)     # [B, child] parent bitmask
    keys = masks * 16 + np.arange(11)[None, :]
    uniq, inv = np.unique(keys, return_inverse=True)
    fams = [(int(k % 16), [p for p in range(11) if (k // 16) >> p & 1]) for k in uniq]
    node, off, par = csr_of(fams)
    fam_scores = C.score_families(codes, card, node, off, par)
    want = fam_scores[inv.reshape(keys.shape)]
    want = np.cumsum(want, axis=1)[:, -1]          # sequential sum in variable order, like the GPU
    assert_scores(got, want)
    assert sachs_scorer.cache_stats()["families"] == len(uniq)
    # adjacency entry point agrees bit for bit
    assert np.array_equal(sachs_scorer.score_adjacency(adj[:5000]), got[:5000])


def test_cyclic_and_invalid_dags(asia_scorer):
    adj = np.zeros((5, 8, 8), dtype=np.uint8)
    adj[0, 0, 1] = adj[0, 1, 2] = 1                 # fine
    adj[1, 0, 1] = adj[1, 1, 2] = adj[1, 2, 0] = 1  # 3-cycle
    adj[2, 3, 3] = 1                                # self loop
    adj[3, 6, 7] = adj[3, 7, 6] = 1                 # 2-cycle
    out, invalid = asia_scorer.score_adjacency(adj, return_invalid=True)
    assert invalid == 3
    assert np.isnan(out[[1, 2, 3]]).all() and not np.isnan(out[[0, 4]]).any()
    labels = np.tile(np.arange(8, dtype=np.uint8), (2, 1))
    labels[1, 7] = 6                                # not a permutation (bnlearn.py:35)
    out, invalid = asia_scorer.score_wire(labels, np.zeros((2, 8), dtype=np.uint32), return_invalid=True)
    assert invalid == 1 and np.isnan(out[1]) and not np.isnan(out[0])
    with pytest.raises(pkg.BicError, match="BAD_FAMILY"):
        asia_scorer.score_families([0], [[0]])
    with pytest.raises(pkg.BicError, match="BAD_FAMILY"):
        asia_scorer.score_families([0], [[8]])


def test_empty_batches(asia_scorer):
    assert asia_scorer.score_adjacency(np.zeros((0, 8, 8), dtype=np.uint8)).shape == (0,)
    assert asia_scorer.score_families([], []).shape == (0,)
    assert asia_scorer.count_families([], []) == []


def test_csr_entry_point_matches_adjacency(sachs_scorer):
    adj = synth.er_candidates(11, 3000, 10, 22, None, seed=9)
    parents, off = [], [0]
    for b in range(adj.shape[0]):
        for i in range(11):
            ps = np.flatnonzero(adj[b, :, i])
            parents.extend(ps.tolist())
            off.append(len(parents))
    got = sachs_scorer.score_csr(np.array(off), np.array(parents, dtype=np.int32), adj.shape[0])
    assert np.array_equal(got, sachs_scorer.score_adjacency(adj))


def test_device_pointer_entry(asia, asia_scorer):
    import torch
    adj = synth.er_candidates(8, 512, 7, 11, None, seed=2)
    host = asia_scorer.score_adjacency(adj)
    dev = asia_scorer.score_adjacency(torch.from_numpy(adj).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)
    with pkg.BicScorer(torch.from_numpy(asia[0]).cuda(), asia[1]) as s2:
        assert np.array_equal(s2.score_adjacency(adj), host)


# ------------------------------------------------------------------------------ edge cases
@pytest.mark.parametrize("N", [1, 15, 16, 17, 127, 129, 4097])
def test_ragged_row_counts(N):
    rng = np.random.default_rng(N)
    card = np.array([2, 3, 1, 4, 5], dtype=np.int32)          # includes a constant variable
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    fams = all_families(5)
    with pkg.BicScorer(codes, card) as s:
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
        for (i, ps), t in zip(fams, tabs):
            assert np.array_equal(t, O.family_counts(codes, card, i, ps)), (N, i, ps)
        node, off, par = csr_of(fams)
        assert_scores(s.score_families_csr(node, off, par), C.score_families(codes, card, node, off, par))


def test_bad_dataset_rejected():
    codes = np.array([[0, 1, 2, 1]], dtype=np.uint8)
    with pytest.raises(pkg.BicError, match="BAD_CODE"):
        pkg.BicScorer(codes, np.array([2], dtype=np.int32))
    with pytest.raises(pkg.BicError, match="BIC_ERR_ARG"):
        pkg.BicScorer(codes, np.array([0], dtype=np.int32))


def test_table_too_large_is_an_error():
    rng = np.random.default_rng(0)
    card = np.full(12, 6, dtype=np.int32)
    codes = rng.integers(0, 6, size=(12, 1000)).astype(np.uint8)
    with pkg.BicScorer(codes, card) as s:
        with pytest.raises(pkg.BicError, match="TABLE_TOO_LARGE"):
            s.score_families([0], [list(range(1, 12))])       # 6^12 cells
        assert s.cache_stats()["families"] == 0
        assert_scores(s.score_families([0], [[1, 2]]), [O.family_score(codes, card, 0, [1, 2])])


def test_wide_network_csr_keys(tmp_path):
    """n = 441 (pigs-shaped): 7-word parent bitmasks, CSR input, in-degree <= 2."""
    n, N = 441, 20000
    adj, card, cpts = synth.make_network(n, 592, 2, [3], seed=11)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(3))
    dags = [adj]
    rng = np.random.default_rng(4)
    for _ in range(3):
        a = adj.copy()
        for _ in range(10):
            u, v = rng.choice(n, 2, replace=False)
            a[u, v] = 0
        dags.append(a)
    parents, off = [], [0]
    for a in dags:
        for i in range(n):
            parents.extend(np.flatnonzero(a[:, i]).tolist())
            off.append(len(parents))
    with pkg.BicScorer(codes, card) as s:
        got = s.score_csr(np.array(off), np.array(parents, dtype=np.int32), len(dags))
        want = C.score_dags_adj(codes, card, np.stack(dags))
        assert_scores(got, want)
        assert np.array_equal(s.score_adjacency(np.stack(dags)), got)
        cyc = adj.copy()
        order = synth.topo_order(adj)
        cyc[order[-1], order[0]] = 1
        ps = np.flatnonzero(adj[:, order[-1]])
        if len(ps):
            cyc[order[0], ps[0]] = 1   # make sure first..last are connected: first -> ps[0] -> last -> first
        out = s.score_adjacency(cyc[None])
        assert np.isnan(out[0]) == (not O.is_acyclic(cyc))


# ------------------------------------------------------------- sliced rows (large N) + props
def test_large_N_sliced_counts_and_properties():
    """N = 3 M rows: few families per launch => several row slices per family, merged through
    HBM tables and reduced by the last slice.  Checked against the C oracle and through
    size-independent properties (sum of counts, marginalisation, decomposition)."""
    N = 3_000_000
    adj, card, cpts = synth.make_network(10, 14, 3, [2, 3, 4], seed=5)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(6))
    fams = [(0, []), (1, [0]), (2, [0, 1]), (3, [0, 1, 2]), (4, [0, 1, 2, 3]), (5, [0, 1, 2, 3, 4]),
            (6, [0, 1, 2, 3, 4, 5]), (7, [0, 1, 2, 3, 4, 5, 6]), (9, [1, 2, 3, 4, 5, 6, 7, 8])]
    with pkg.BicScorer(codes, card) as s:
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
        for (i, ps), t in zip(fams, tabs):
            assert t.sum() == N
            assert np.array_equal(t, C.family_counts(codes, card, i, ps)), (i, ps)
        # marginalising the last parent out gives the smaller family's table
        t_big = s.family_counts(4, [0, 1, 2, 3])
        t_small = s.family_counts(4, [0, 1, 2])
        r3 = int(card[3])
        assert np.array_equal(t_big.reshape(-1, r3, int(card[4])).sum(axis=1), t_small)
        node, off, par = csr_of(fams)
        got = s.score_families_csr(node, off, par, no_cache=True)
        assert_scores(got, C.score_families(codes, card, node, off, par))
        # a DAG's score is the sum of its family terms, in variable order
        dag = np.zeros((1, 10, 10), dtype=np.uint8)
        dag[0] = adj
        fam_terms = s.score_families(list(range(10)), [np.flatnonzero(adj[:, i]).tolist() for i in range(10)])
        assert s.score_adjacency(dag)[0] == np.cumsum(fam_terms)[-1]
        # many families in one launch (one slice each) give the same bits as sliced launches
        many = all_families(10, max_k=2)
        node, off, par = csr_of(many)
        batch = s.score_families_csr(node, off, par, no_cache=True)
        s.cache_clear()
        single = np.array([s.score_families([i], [ps], no_cache=True)[0] for i, ps in many[:40]])
        assert np.array_equal(batch[:40], single)


# ------------------------------------------------------ BASELINE.json full size (config 4)
def test_alarm_shaped_full_size_10M_rows():
    """configs[3]: 37 variables, 10 M rows (the bench workload).  Oracle on a handful of families
    and DAGs (seconds), plus size-independent properties over a whole candidate batch."""
    import torch
    import bench
    cfg = bench.WORKLOADS["alarm"]
    N = cfg["rows"]
    true_adj, card, codes_t = bench.make_dataset_gpu(cfg, N, torch.device("cuda", 0))
    codes = codes_t.cpu().numpy()
    with pkg.BicScorer(codes_t, card) as s:
        del codes_t
        rng = np.random.default_rng(0)
        fams = []
        for k in (0, 1, 2, 3, 4, 5, 6):
            i = int(rng.integers(37))
            fams.append((i, sorted(rng.choice([p for p in range(37) if p != i], size=k, replace=False).tolist())))
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
        for (i, ps), t in zip(fams, tabs):
            assert t.sum() == N
            assert np.array_equal(t, C.family_counts(codes, card, i, ps)), (i, ps)
        # marginalisation: dropping the last parent of the k=4 family
        i, ps = fams[4]
        small = s.family_counts(i, ps[:-1])
        assert np.array_equal(tabs[4].reshape(-1, int(card[ps[-1]]), int(card[i])).sum(axis=1), small)
        # a batch of candidates: cold == warm == no-cache bits; true DAG beats its neighbours' mean
        cand = bench.candidate_batch(cfg, 512, 0, 0, 1)
        cold = s.score_adjacency(cand, no_cache=True)
        warm = s.score_adjacency(cand)
        assert np.array_equal(cold, warm) and not np.isnan(cold).any()
        want = C.score_dags_adj(codes, card, np.concatenate([true_adj[None], cand[:3]]))
        got = s.score_adjacency(np.concatenate([true_adj[None], cand[:3]]))
        assert_scores(got, want)
        assert got[0] > cold.mean()
        # decomposition: DAG score == sum of its family terms in variable order
        terms = s.score_families(list(range(37)), [np.flatnonzero(cand[7][:, v]).tolist() for v in range(37)])
        assert cold[7] == np.cumsum(terms)[-1]


# ------------------------------------------- tables marginalised from counted supersets
def test_derived_families_match_counted_ones():
    """On large datasets a new family whose superset (one more parent) is counted in the same
    batch gets its table by summing the superset's table over that parent.  Integer sums: the
    scores must be bit-identical to counting every family from the rows, and match the oracle."""
    N = 1_200_000
    adj, card, cpts = synth.make_network(9, 12, 3, [2, 3, 4, 1], seed=21)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(22))
    # nested parent sets of node 0 and 5 -> chains of donors, plus unrelated families
    fams = [(0, []), (0, [1]), (0, [1, 2]), (0, [1, 2, 3]), (0, [1, 2, 3, 4]), (0, [2, 3]), (0, [3]),
            (5, [0, 8]), (5, [0, 6, 8]), (5, [0, 1, 6, 8]), (5, [8]), (7, [6]), (7, [2, 6]), (4, [0, 1, 2, 3, 5, 6]),
            # donors with a different child: {2,4,6} is inside {1,2,4,6} = family (1, [2,4,6]); {3,8} inside (8, [2,3])
            (1, [2, 4, 6]), (6, [2, 4]), (4, [2, 6]), (8, [2, 3]), (3, [8]), (2, [3]), (3, []), (8, [])]
    node, off, par = csr_of(fams)
    with pkg.BicScorer(codes, card) as s:
        s.profile_reset()
        derived = s.score_families_csr(node, off, par, no_cache=True)
        prof = s.profile()
        assert prof["families_derived"] >= 9 and prof["families_counted"] + prof["families_derived"] == len(fams)
        s.derive = False
        s.profile_reset()
        counted = s.score_families_csr(node, off, par, no_cache=True)
        assert s.profile()["families_derived"] == 0
        assert np.array_equal(derived, counted)
        assert_scores(derived, C.score_families(codes, card, node, off, par))
        # DAG batches: random candidates, derive on/off identical bits
        s.derive = True
        dags = synth.er_candidates(9, 600, 8, 20, 5, seed=23)
        a = s.score_adjacency(dags, no_cache=True)
        s.derive = False
        b = s.score_adjacency(dags, no_cache=True)
        assert np.array_equal(a, b)
        assert_scores(a[:40], C.score_dags_adj(codes, card, dags[:40]))


# ------------------------------------------------- 2-bit packed shadow copy of the dataset
@pytest.mark.parametrize("two", ["0", "1"])
@pytest.mark.parametrize("N", [1, 63, 64, 65, 511, 513, 4097, 70001])
def test_packed_path_ragged_rows(N, two, monkeypatch):
    """Columns with <= 4 states are also held 4 rows per byte and streamed from there.  Force
    that path on small, ragged row counts (tail masking of the 64-row groups) and on families
    with 1..7 columns (single- and two-group index arithmetic); mixed with a 5-state column
    that must fall back to the uint8 path."""
    monkeypatch.setenv("BIC_PACK2_MIN_ROWS", "1")
    monkeypatch.setenv("BIC_P2_TWO", two)                    # families of <= 3 columns: two 64-row groups in flight
    rng = np.random.default_rng(N)
    card = np.array([2, 3, 4, 4, 3, 2, 4, 5, 1, 3], dtype=np.int32)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    fams = [(0, []), (1, [0]), (2, [0, 1]), (3, [0, 1, 2]), (4, [0, 1, 2, 3]), (5, [0, 1, 2, 3, 4]),
            (6, [0, 1, 2, 3, 4, 5]), (9, [0, 1, 2, 3, 4, 5, 6]), (6, [1, 3, 7]), (7, [2, 3]), (3, [8, 9]),
            (2, [3, 4, 6, 9]), (0, [1, 2, 3, 4, 5, 6, 9])]
    with pkg.BicScorer(codes, card) as s:
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
        for (i, ps), t in zip(fams, tabs):
            assert np.array_equal(t, O.family_counts(codes, card, i, ps)), (N, i, ps)
        node, off, par = csr_of(fams)
        assert_scores(s.score_families_csr(node, off, par, no_cache=True), C.score_families(codes, card, node, off, par))


@pytest.mark.parametrize("N", [70_001, 1_100_003])
def test_packed_path_offsets_replicas_swizzle(N, monkeypatch):
    """The last stage of the packed path (csrc/count_kernels.cuh, p2_group): counter offsets by
    integer dot product, with and without a high column group, and the 16-bit-lane fallback when
    the four low columns have four states each (their product does not fit a byte weight); lane
    replicas beyond cells * R = 16383 and the 512-thread x 96 KB class-0 shape (class0_shape);
    the bank swizzle of un-replicated tables (few rows: every table; many rows: tables above the
    replica reach and class 1); the class-0 / class-1 lists counted by two launches with different
    CTA shapes (tiers).  Counts equal the oracle's; the knobs that turn each piece off and the uint8
    path give the same bits."""
    monkeypatch.setenv("BIC_PACK2_MIN_ROWS", "1")
    rng = np.random.default_rng(N)
    card = np.array([4, 4, 4, 4, 4, 4, 3, 2, 4, 3, 2, 2], dtype=np.int32)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    codes[4] = (codes[0] + codes[1] * (codes[2] > 1)) % 4                    # structure: skewed cells
    codes[11] = (codes[3] > 0).astype(np.uint8) & codes[10]
    fams = [(4, [0, 1, 2, 3]),          # low columns 4 x 4 x 4 x 4 = 256: 16-bit-lane fallback; 1024 cells
            (5, [0, 1, 2, 3, 4]),       # the same with two high columns... one: 4096 cells, class 1, fallback
            (6, [0, 1, 2, 3, 4, 5]),    # 12 288 cells, class 1, low product 192: dot product with a high group
            (7, [0, 1, 2]),             # 128 cells, no high group: one IDP per row
            (9, [0, 1, 6, 7]),          # 288 cells
            (10, [0, 1, 2, 8]),         # 512 cells: 32 replicas only in the wide shape
            (11, [0, 1, 2, 3, 6]),      # 1536 cells: 16 replicas only without the 16-bit cap
            (11, [0, 1, 2, 3, 4]),      # 2048 cells: above the replica reach, swizzled
            (8, [0, 1, 2, 3, 6, 7]),    # 6144 cells, class 1
            (3, []), (2, [3])]
    node, off, par = csr_of(fams)
    want_tabs = [C.family_counts(codes, card, i, ps) for i, ps in fams]

    def run():
        with pkg.BicScorer(codes, card) as s:
            tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
            for (i, ps), t, w in zip(fams, tabs, want_tabs):
                assert np.array_equal(t, w), (N, i, ps)
            return s.score_families_csr(node, off, par, no_cache=True)

    base = run()
    assert_scores(base, C.score_families(codes, card, node, off, par))
    # BIC_TIER0 / BIC_TIER1 = 0: one launch per class list instead of two with different CTA shapes
    for knob in ("BIC_SWIZZLE=0", "BIC_CLASS0_WIDE=0", "BIC_TIER0=0", "BIC_TIER1=0", "BIC_NO_PACK2=1"):
        k, v = knob.split("=")
        monkeypatch.setenv(k, v)
        assert np.array_equal(run(), base), knob
        monkeypatch.delenv(k)


def test_packed_and_byte_paths_agree(monkeypatch):
    N = 600_000
    adj, card, cpts = synth.make_network(12, 18, 3, [2, 3, 4], seed=41)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(42))
    dags = synth.er_candidates(12, 400, 11, 30, 6, seed=43)
    with pkg.BicScorer(codes, card) as s:
        packed = s.score_adjacency(dags, no_cache=True)
    monkeypatch.setenv("BIC_NO_PACK2", "1")
    with pkg.BicScorer(codes, card) as s:
        plain = s.score_adjacency(dags, no_cache=True)
    assert np.array_equal(packed, plain)
    assert_scores(packed[:30], C.score_dags_adj(codes, card, dags[:30]))


# ----------------------------------- tables above one CTA's shared memory (class 3), slicing
def test_class3_cluster_passes_and_l2_atomics_match_oracle(monkeypatch):
    """Tables of more than 49152 cells.  When the rows dwarf the table it is counted in passes over
    shared-memory sub-ranges (k_count<1024,false,true>, the default) or in ONE pass by a
    thread-block cluster whose CTAs share the table in distributed shared memory (k_count_cluster,
    BIC_CLUSTER=1; measured slower); otherwise straight into HBM with L2 atomics.  All three must give the oracle's
    counts and identical score bits; ragged row count, k <= 6 (specialised row loop) and k > 6
    (generic loop), one and several row slices, cluster sizes 2 / 4 / 8."""
    N = 1_200_003
    rng = np.random.default_rng(77)
    card = np.array([21, 20, 19, 5, 7] + [3] * 11, dtype=np.int32)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    codes[3] = (codes[0] + codes[4]) % 5                      # structure: not every cell is hit equally
    fams = [(4, [0, 1, 2]),                                   # 55 860 cells: 2 CTAs / 2 passes
            (0, [1, 2, 3, 4]),                                # 279 300 cells: 8 CTAs / 6 passes
            (5, list(range(6, 16))),                          # 3^11 = 177 147 cells, k = 10
            (3, [0, 1, 2])]                                   # 39 900 cells: class 2, same launch sequence
    node, off, par = csr_of(fams)
    want_tabs = [C.family_counts(codes, card, i, ps) for i, ps in fams]
    want_scores = C.score_families(codes, card, node, off, par)

    def run():
        with pkg.BicScorer(codes, card) as s:
            tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
            for (i, ps), t, w in zip(fams, tabs, want_tabs):
                assert t.sum() == N
                assert np.array_equal(t, w), (i, ps)
            scores = s.score_families_csr(node, off, par, no_cache=True)
            one = s.score_families([4], [[0, 1, 2]], no_cache=True)          # alone: several row slices, smaller cluster
            assert one[0] == scores[0]
            assert np.array_equal(s.family_counts(4, [0, 1, 2]), tabs[0])
            two = s.score_families([5], [list(range(6, 16))], no_cache=True)     # 4 CTAs per cluster when alone
            assert two[0] == scores[2]
        assert_scores(scores, want_scores)
        return scores

    ranged = run()                                            # sub-range passes
    monkeypatch.setenv("BIC_PARK_CELLS", "1")
    assert np.array_equal(run(), ranged)                      # passes on cell indices parked once by k_cells
    monkeypatch.setenv("BIC_CELLS_MAX_MB", "1")
    assert np.array_equal(run(), ranged)                      # scratch limit exceeded: recompute
    monkeypatch.delenv("BIC_CELLS_MAX_MB")
    monkeypatch.delenv("BIC_PARK_CELLS")
    monkeypatch.setenv("BIC_CLUSTER", "1")
    assert np.array_equal(run(), ranged)                      # thread-block clusters
    monkeypatch.setenv("BIC_CLUSTER_THREADS", "512")
    assert np.array_equal(run(), ranged)
    monkeypatch.delenv("BIC_CLUSTER_THREADS")
    monkeypatch.setenv("BIC_CLUSTER_SIZE", "8")
    assert np.array_equal(run(), ranged)
    monkeypatch.delenv("BIC_CLUSTER_SIZE")
    monkeypatch.delenv("BIC_CLUSTER")
    monkeypatch.setenv("BIC_RANGE_PASSES", "0")
    assert np.array_equal(run(), ranged)                      # L2 atomics
    monkeypatch.delenv("BIC_RANGE_PASSES")
    monkeypatch.setenv("BIC_C3_U16", "1")
    assert np.array_equal(run(), ranged)                      # 16-bit counters: half the passes, spills between phases


def test_class3_top_split_matches_generic_cut(monkeypatch):
    """Class-3 sub-ranges along the states of the first parent (range_plan in csrc/common.cuh: the
    pass tests that parent's byte alone and builds the rest of the index on 16-bit packed lanes)
    against the generic cut by cell index (BIC_TOPSPLIT=0): k = 1 .. 6 parents, a last pass with
    fewer states, a family whose table below the first parent is too large for the split, ragged
    row count.  Counts must equal the oracle's, scores must be the same bits either way."""
    N = 1_300_003
    rng = np.random.default_rng(78)
    card = np.array([3, 250, 240, 21, 20, 6, 5, 4, 2], dtype=np.int32)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    codes[5] = (codes[3] + codes[4]) % 6                      # structure: not every cell is hit equally
    codes[1] = np.minimum(codes[1], rng.integers(0, 250, size=N)).astype(np.uint8)   # skewed first parent
    fams = [(2, [1]),                                         # k = 1: 250 x 240 = 60 000 cells, 2 passes of 125 states
            (3, [1, 4]),                                      # k = 2: 250 x 20 x 21 = 105 000 cells, 3 passes of 84 states
            (8, [1, 3, 4]),                                   # k = 3: 250 x 21 x 20 x 2 = 210 000 cells, 5 passes
            (7, [3, 4, 5, 6]),                                # k = 4: 21 x 20 x 6 x 5 x 4 = 50 400 cells, passes of 11 and 10 states
            (7, [3, 4, 5, 6, 8]),                             # k = 5: 201 600 cells, 5 passes of 5 states (the last holds one)
            (0, [3, 4, 5, 6, 7, 8]),                          # k = 6: 302 400 cells, 7 passes of 3 states
            (8, [0, 3, 4, 5, 6, 7]),                          # k = 6, first parent has 3 states, 100 800 cells below it: no split
            (4, [0, 2, 3])]                                   # k = 3, 3 x 240 x 21 x 20 = 302 400 cells: no split either
    node, off, par = csr_of(fams)
    want_tabs = [C.family_counts(codes, card, i, ps) for i, ps in fams]
    want_scores = C.score_families(codes, card, node, off, par)

    def run():
        with pkg.BicScorer(codes, card) as s:
            tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
            for (i, ps), t, w in zip(fams, tabs, want_tabs):
                assert t.sum() == N
                assert np.array_equal(t, w), (i, ps)
            scores = s.score_families_csr(node, off, par, no_cache=True)
            alone = s.score_families([7], [[3, 4, 5, 6, 8]], no_cache=True)      # alone: several row slices
            assert alone[0] == scores[4]
        assert_scores(scores, want_scores)
        return scores

    split = run()
    monkeypatch.setenv("BIC_TOPSPLIT", "0")
    assert np.array_equal(run(), split)
    monkeypatch.delenv("BIC_TOPSPLIT")
    monkeypatch.setenv("BIC_CLASS2_THREADS", "512")          # k_count<512, false, true>
    assert np.array_equal(run(), split)
    monkeypatch.setenv("BIC_C3_U16", "1")                    # 16-bit counters, phase spills (count_rows_r16), 512 threads
    assert np.array_equal(run(), split)
    monkeypatch.delenv("BIC_CLASS2_THREADS")
    assert np.array_equal(run(), split)                      # ... and 1024 threads


def test_slice_choice_does_not_change_bits(monkeypatch):
    """The number of row slices per family is a cost decision (L2 windows vs merge traffic);
    counts are integer sums and the fp64 reduce has a fixed order, so any choice gives the same
    bits."""
    N = 1_300_000
    adj, card, cpts = synth.make_network(40, 60, 3, [2, 3, 5, 9], seed=51)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(52))     # 52 MB: two L2 windows
    dags = synth.er_candidates(40, 6, 39, 70, 4, seed=53)
    with pkg.BicScorer(codes, card) as s:
        a = s.score_adjacency(dags, no_cache=True)
    monkeypatch.setenv("BIC_SLICE_MODEL", "0")
    with pkg.BicScorer(codes, card) as s:
        b = s.score_adjacency(dags, no_cache=True)
    assert np.array_equal(a, b)
    assert_scores(a[:2], C.score_dags_adj(codes, card, dags[:2]))


# ------------------------------------------------------- sub-batching, cache growth, misc API
def test_large_batch_spans_sub_batches(asia, asia_scorer):
    """600k DAGs x 8 nodes = 4.8 M instances > the 4 M-instance sub-batch: two passes through the
    dedup / count / gather pipeline, host CSR offsets re-based per sub-batch."""
    codes, card = asia
    base = synth.er_candidates(8, 3000, 5, 14, None, seed=77)
    reps = 200
    adj = np.concatenate([base] * reps)
    want = np.tile(C.score_dags_adj(codes, card, base), reps)
    asia_scorer.cache_clear()
    got = asia_scorer.score_adjacency(adj)
    assert got.shape == (600000,)
    assert_scores(got, want)
    # CSR entry point over the same batch
    b, c_, p = np.nonzero(adj.transpose(0, 2, 1))
    counts = np.bincount(b * 8 + c_, minlength=adj.shape[0] * 8)
    off = np.zeros(adj.shape[0] * 8 + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    got_csr = asia_scorer.score_csr(off, p.astype(np.int32), adj.shape[0])
    assert np.array_equal(got_csr, got)
    import torch
    dev = asia_scorer.score_adjacency(torch.from_numpy(adj).cuda())
    assert np.array_equal(dev.cpu().numpy(), got)


def test_cache_growth_reserve_and_stats(sachs):
    codes, card = sachs
    with pkg.BicScorer(codes, card) as s:
        st0 = s.cache_stats()
        assert st0["families"] == 0
        s.cache_reserve(3_000_000)                 # beyond the 2^20-family initial size: grows
        assert s.cache_stats()["capacity"] >= 3_000_000
        fams = all_families(11, max_k=3)          # 11 * (1 + 10 + 45 + 120) = 1936 families
        node, off, par = csr_of(fams)
        a = s.score_families_csr(node, off, par)
        st = s.cache_stats()
        assert st["families"] == len(fams) and st["misses"] == len(fams) and st["lookups"] == len(fams)
        b = s.score_families_csr(node, off, par)     # all hits
        st = s.cache_stats()
        assert st["misses"] == len(fams) and st["lookups"] == 2 * len(fams)
        assert np.array_equal(a, b)
        # growth by rehash keeps every cached family
        dags = synth.er_candidates(11, 150_000, 10, 25, None, seed=5)
        s.score_adjacency(dags)
        assert np.array_equal(s.score_families_csr(node, off, par), a)
        s.cache_clear()
        assert s.cache_stats()["families"] == 0
        assert np.array_equal(s.score_families_csr(node, off, par), a)


def test_stream_profile_and_metric_switch(asia, asia_scorer):
    import torch
    codes, card = asia
    adj = synth.er_candidates(8, 256, 7, 12, None, seed=3)
    ref = asia_scorer.score_adjacency(adj, no_cache=True)
    stream = torch.cuda.Stream()
    asia_scorer.set_stream(stream.cuda_stream)
    asia_scorer.profile_enable(True)
    asia_scorer.profile_reset()
    got = asia_scorer.score_adjacency(adj, no_cache=True)
    prof = asia_scorer.profile()
    asia_scorer.profile_enable(False)
    asia_scorer.set_stream(None)
    assert np.array_equal(got, ref)
    assert prof["count_launches"] >= 1 and prof["kernel_launches"] > prof["count_launches"]
    assert prof["families_counted"] + prof["families_derived"] == asia_scorer.cache_stats()["families"]
    assert prof["count_ms"] > 0 and prof["alg_bytes"] > 0 and prof["rows_counted"] == prof["families_counted"] * 5000
    # one cache serves every metric: penalties are applied at gather time
    ll = asia_scorer.score_adjacency(adj, metric="loglik")
    aic = asia_scorer.score_adjacency(adj, metric="aic")
    misses = asia_scorer.cache_stats()["misses"]
    bic = asia_scorer.score_adjacency(adj, metric="bic")
    assert asia_scorer.cache_stats()["misses"] == misses
    nparams = ll - aic
    assert np.allclose(bic, ll - 0.5 * np.log(5000) * nparams, rtol=1e-13)
    with pytest.raises(NotImplementedError):
        asia_scorer.score_adjacency(adj, metric="mbde")


# ------------------------------------------------------------------ bde (BDeu) and k2
def test_bde_and_k2_match_oracle(asia, sachs):
    """bnlearn's other decomposable count-based scores, computed as an lgamma epilogue on the same
    count tables.  The reference pins none of them: oracle-only parity."""
    codes, card = asia
    fams = all_families(8)
    node, off, par = csr_of(fams)
    with pkg.BicScorer(codes, card) as s:
        for metric, iss in (("bde", 1.0), ("bde", 10.0), ("k2", 1.0)):
            s.set_iss(iss)
            C.set_iss(iss)
            got = s.score_families_csr(node, off, par, metric=metric)
            assert_scores(got, C.score_families(codes, card, node, off, par, metric=metric))
            assert got[5] == pytest.approx(O.family_score(codes, card, fams[5][0], fams[5][1], metric, iss), rel=1e-9)
        C.set_iss(1.0)
        s.set_iss(1.0)
        # switching between term kinds re-counts (one kind cached at a time) and stays consistent
        adj = synth.er_candidates(8, 500, 5, 14, None, seed=12)
        bde = s.score_adjacency(adj, metric="bde")
        bic = s.score_adjacency(adj, metric="bic")
        assert np.array_equal(s.score_adjacency(adj, metric="bde"), bde)
        assert_scores(bic, C.score_dags_adj(codes, card, adj, metric="bic"))
        assert_scores(bde, C.score_dags_adj(codes, card, adj, metric="bde"))
    ev = pkg.BNLearnWrapper("asia", "k2")
    g = Graph(8, [(0, 2), (1, 3), (1, 4), (2, 5), (3, 5), (5, 6), (5, 7), (4, 7)], list(range(8)))
    true = np.zeros((1, 8, 8), dtype=np.uint8)
    for u, v in g.get_edgelist():
        true[0, u, v] = 1
    assert ev.score(g) == pytest.approx(C.score_dags_adj(codes, card, true, metric="k2")[0], rel=1e-9)
    # large tables (HBM class) and a big dataset through the derive path
    codes, card = sachs
    with pkg.BicScorer(codes, card, metric="bde", iss=5.0) as s:
        C.set_iss(5.0)
        f2 = [(0, list(range(1, 11))), (3, [0, 1, 2, 4, 5, 6, 7, 8]), (4, [1])]
        node, off, par = csr_of(f2)
        assert_scores(s.score_families_csr(node, off, par), C.score_families(codes, card, node, off, par, metric="bde"))
        C.set_iss(1.0)
    N = 1_100_000
    adj, card, cpts = synth.make_network(8, 10, 3, [2, 3], seed=51)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(52))
    dags = synth.er_candidates(8, 200, 7, 16, 5, seed=53)
    with pkg.BicScorer(codes, card, metric="k2") as s:
        got = s.score_adjacency(dags)
        assert s.profile()["families_derived"] > 0
        assert_scores(got[:25], C.score_dags_adj(codes, card, dags[:25], metric="k2"))


# ------------------------------------------------------- search loop on top of the scorer
def test_hill_climb_search_loop(asia, asia_scorer):
    """Config 3's "search loop with batched scoring": greedy hill climbing over single-edge moves.
    Every iteration is one batched call; the search must be monotone, end in a local optimum that
    the oracle confirms, beat the empty graph by a wide margin, and run almost entirely out of the
    family-score cache.  (The path itself is not compared with a CPU run: BIC is score-equivalent,
    so exact ties between edge orientations are broken by rounding.)"""
    from dags_vae_search_b200 import search
    codes, card = asia
    asia_scorer.cache_clear()
    best, score, trace = search.hill_climb(asia_scorer)
    assert O.is_acyclic(best) and len(trace) >= 5
    steps = [s for _, s in trace]
    assert all(b > a for a, b in zip(steps, steps[1:]))
    assert score == pytest.approx(O.score_adjacency(codes, card, best), rel=RTOL)
    empty = O.score_adjacency(codes, card, np.zeros((8, 8), dtype=np.uint8))
    assert score > empty + 1000 and score > -11200          # the true DAG scores -11109.74 on this sample
    cand, _ = search.neighbours(best)
    assert all(O.is_acyclic(a) for a in cand[:50])
    assert (asia_scorer.score_adjacency(cand) <= score + 1e-6).all()      # local optimum
    st = asia_scorer.cache_stats()
    assert st["misses"] <= 1024 and st["lookups"] > 20 * st["misses"]
    # wire format straight from CUDA tensors (decoder output that never leaves the GPU)
    import torch
    labels = np.tile(np.arange(8, dtype=np.uint8), (3, 1))
    ebits = np.zeros((3, 8), dtype=np.uint32)
    ebits[1, 2] = 0b11
    ebits[2, 7] = 0b1010101
    host = asia_scorer.score_wire(labels, ebits)
    dev = asia_scorer.score_wire(torch.from_numpy(labels).cuda(), torch.from_numpy(ebits.astype(np.int32)).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), host)


def test_degenerate_shapes():
    """One variable, one row, all-constant columns, a single DAG."""
    with pkg.BicScorer(np.array([[0, 1, 1, 0, 1]], dtype=np.uint8), np.array([2], dtype=np.int32)) as s:
        got = s.score_adjacency(np.zeros((1, 1, 1), dtype=np.uint8))
        want = 2 * np.log(2 / 5) + 3 * np.log(3 / 5) - 0.5 * np.log(5)
        assert got[0] == pytest.approx(want, rel=1e-12)
        out, bad = s.score_adjacency(np.ones((1, 1, 1), dtype=np.uint8), return_invalid=True)   # self loop
        assert bad == 1 and np.isnan(out[0])
    codes = np.zeros((3, 1), dtype=np.uint8)
    with pkg.BicScorer(codes, np.array([1, 1, 1], dtype=np.int32)) as s:       # one row, constant variables
        adj = np.zeros((1, 3, 3), dtype=np.uint8)
        adj[0, 0, 1] = adj[0, 1, 2] = 1
        assert s.score_adjacency(adj)[0] == 0.0
        assert s.family_counts(2, [0, 1]).tolist() == [[1]]
    with pkg.BicScorer(codes, np.array([2, 3, 4], dtype=np.int32)) as s:       # one row, declared wider
        adj = np.zeros((1, 3, 3), dtype=np.uint8)
        adj[0, 0, 2] = adj[0, 1, 2] = 1
        # ln N = 0 with one row: BIC == loglik == 0; AIC charges the declared parameters
        assert s.score_adjacency(adj)[0] == 0.0
        assert s.score_adjacency(adj, metric="aic")[0] == -(1 + 2 + 3 * 2 * 3)


def test_small_warm_batches_take_the_one_launch_path(monkeypatch):
    """Small batches whose families are all cached are scored by one kernel (k_score_small: keys,
    cycle check, lookup, sum).  Same bits as the general pipeline, same NaN / n_invalid for cyclic
    DAGs, and a batch with an unseen family falls back (and is then cached)."""
    for n, N in ((8, 3000), (37, 5000), (64, 2000)):
        rng = np.random.default_rng(n)
        card = rng.choice(np.array([2, 3], dtype=np.int32), size=n)
        codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
        dags = synth.er_candidates(n, 40, n - 1, 2 * n, 3, seed=5)
        cyc = dags[3].copy()
        order = synth.topo_order(cyc)
        cyc[order[-1], order[0]] = 1
        ps = np.flatnonzero(dags[3][:, order[-1]])
        if len(ps):
            cyc[order[0], ps[0]] = 1
        loop = dags[4].copy()
        loop[2, 2] = 1
        batch = np.concatenate([dags, cyc[None], loop[None]])
        monkeypatch.setenv("BIC_NO_FAST_SMALL", "1")
        with pkg.BicScorer(codes, card) as s:
            want, want_bad = s.score_adjacency(batch, return_invalid=True)
            want_aic = s.score_adjacency(batch, metric="aic")
            launches_general = s.profile()["kernel_launches"]
        monkeypatch.delenv("BIC_NO_FAST_SMALL")
        with pkg.BicScorer(codes, card) as s:
            cold = s.score_adjacency(batch)                       # cold cache: general pipeline
            assert np.array_equal(cold, want, equal_nan=True)
            assert np.array_equal(s.score_adjacency(batch), want, equal_nan=True)   # all cached now, but the last call had misses: general once more
            before = s.profile()["kernel_launches"]
            warm, bad = s.score_adjacency(batch, return_invalid=True)
            assert s.profile()["kernel_launches"] == before + 1   # one launch
            assert np.array_equal(warm, want, equal_nan=True) and bad == want_bad
            assert np.isnan(warm[-1]) and np.isnan(warm[-2]) == (not O.is_acyclic(cyc))
            assert np.array_equal(s.score_adjacency(batch, metric="aic"), want_aic, equal_nan=True)
            one = np.array([s.score_adjacency(batch[b:b + 1])[0] for b in range(10)])   # the reference's usage: one DAG per call
            assert np.array_equal(one, want[:10])
            # an unseen family: falls back, is counted, and the next call is short again
            fresh = synth.er_candidates(n, 5, n - 1, 2 * n, 3, seed=6)
            fams0 = s.cache_stats()["families"]
            got = s.score_adjacency(np.concatenate([batch[:3], fresh]))
            assert s.cache_stats()["families"] > fams0
            assert_scores(got[3:], C.score_dags_adj(codes, card, fresh))
            before = s.profile()["kernel_launches"]
            s.score_adjacency(fresh)          # all cached, but the last call had misses: general pipeline once more
            s.score_adjacency(fresh)
            assert s.profile()["kernel_launches"] > before + 1
            before = s.profile()["kernel_launches"]
            again = s.score_adjacency(fresh)
            assert s.profile()["kernel_launches"] == before + 1 and np.array_equal(again, got[3:])
        assert launches_general > 10


def test_limits_1024_variables_255_states():
    """The documented limits: n = 1024 variables (16-word parent masks, block-per-DAG cycle check,
    warp-per-DAG gather) and 255 states per variable (fp64 reduce staged in pieces of
    floor(cap / 255) parent configurations; 255 x 255 = 65 025 cells: class 3, counted in two
    sub-range passes here)."""
    n, N = 1024, 300_000
    rng = np.random.default_rng(91)
    card = rng.choice(np.array([2, 3, 4], dtype=np.int32), size=n)
    card[[0, 5, 1023]] = 255
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    codes[1023] = (codes[0].astype(np.int32) * 7 + codes[1]) % 255
    fams = [(1023, [0]), (0, [5]), (5, []), (1022, [0, 1023]), (7, [1, 2, 3, 1021])]
    # a chain DAG over all 1024 variables plus the edges above; CSR input
    parents = [[] for _ in range(n)]
    for i in range(1, n):
        parents[i].append(i - 1)
    parents[1023] = [0, 1022]
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(p) for p in parents])
    flat = np.array([p for ps in parents for p in sorted(ps)], dtype=np.int32)
    with pkg.BicScorer(codes, card) as s:
        tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])     # largest table 255 * 255 * r: L2 atomics
        for (i, ps), t in zip(fams, tabs):
            assert np.array_equal(t, C.family_counts(codes, card, i, ps)), (i, ps)
        assert np.array_equal(s.family_counts(1023, [0]), tabs[0])               # alone: two sub-range passes
        node, foff, fpar = csr_of(fams)
        want_f = C.score_families(codes, card, node, foff, fpar)
        assert_scores(s.score_families_csr(node, foff, fpar, no_cache=True), want_f)
        assert_scores(s.score_families([1023], [[0]], no_cache=True), want_f[:1])
        got = s.score_csr(off, flat, 1)
        want = float(np.cumsum(C.score_families(codes, card, np.arange(n, dtype=np.int32), off, flat))[-1])
        assert got[0] == pytest.approx(want, rel=RTOL)
        # close the chain into a cycle: 1023 -> 0
        cyc = [list(p) for p in parents]
        cyc[0] = [1023]
        coff = np.zeros(n + 1, dtype=np.int64)
        coff[1:] = np.cumsum([len(p) for p in cyc])
        cflat = np.array([p for ps in cyc for p in sorted(ps)], dtype=np.int32)
        assert np.isnan(s.score_csr(coff, cflat, 1)[0])


def test_cache_checkpoint_resume(sachs, tmp_path):
    """Checkpoint / resume of a search: the family cache survives a process restart."""
    codes, card = sachs
    dags = synth.er_candidates(11, 5000, 10, 25, None, seed=61)
    path = str(tmp_path / "cache.npz")
    with pkg.BicScorer(codes, card) as s:
        first = s.score_adjacency(dags)
        saved = s.save_cache(path)
        assert saved == s.cache_stats()["families"] > 1000
    with pkg.BicScorer(codes, card) as s2:
        assert s2.load_cache(path) == saved
        again = s2.score_adjacency(dags)
        st = s2.cache_stats()
        assert np.array_equal(again, first) and st["misses"] == 0 and st["families"] == saved
        # new families still get counted and join the restored ones
        more = synth.er_candidates(11, 500, 10, 25, None, seed=62)
        got = s2.score_adjacency(more)
        assert_scores(got[:20], C.score_dags_adj(codes, card, more[:20]))
    with pkg.BicScorer(codes[:, :4000], card) as s3:
        with pytest.raises(ValueError, match="different dataset"):
            s3.load_cache(path)
