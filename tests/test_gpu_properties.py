"""GPU: hypothesis-driven parity on random small problems (ragged N, mixed cardinalities
including constant columns, arbitrary DAGs) and invariants the domain offers."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import dags_vae_search_b200 as pkg
from oracle import bic_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scorer():
    with pkg.BicScorer(np.zeros((1, 1), dtype=np.uint8), np.array([1], dtype=np.int32)) as s:
        yield s


@st.composite
def problems(draw):
    n = draw(st.integers(2, 7))
    N = draw(st.integers(1, 300))
    card = np.array(draw(st.lists(st.integers(1, 5), min_size=n, max_size=n)), dtype=np.int32)
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    codes = np.stack([rng.integers(0, c, size=N) for c in card]).astype(np.uint8)
    B = draw(st.integers(1, 6))
    adj = np.zeros((B, n, n), dtype=np.uint8)
    for b in range(B):
        perm = rng.permutation(n)
        dens = rng.random()
        for i in range(n):
            for u in range(i):
                if rng.random() < dens:
                    adj[b, perm[u], perm[i]] = 1
    return codes, card, adj


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(problems(), st.sampled_from(["bic", "aic", "loglik"]))
def test_random_problems_match_oracle(scorer, problem, metric):
    codes, card, adj = problem
    scorer.set_dataset(codes, card)
    got = scorer.score_adjacency(adj, metric=metric)
    want = np.array([O.score_adjacency(codes, card, a, metric) for a in adj])
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)
    n = codes.shape[0]
    for i in range(n):
        ps = np.flatnonzero(adj[0][:, i]).tolist()
        t = scorer.family_counts(i, ps)
        assert np.array_equal(t, O.family_counts(codes, card, i, ps))
        assert t.sum() == codes.shape[1]


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(problems())
def test_invariants(scorer, problem):
    codes, card, adj = problem
    n, N = codes.shape
    scorer.set_dataset(codes, card)
    base = scorer.score_adjacency(adj)
    # rows are exchangeable
    perm = np.random.default_rng(0).permutation(N)
    scorer.set_dataset(np.ascontiguousarray(codes[:, perm]), card)
    assert np.allclose(scorer.score_adjacency(adj), base, rtol=1e-12, atol=1e-9)
    # renaming the states of a variable leaves every score unchanged
    flipped = codes.copy()
    flipped[0] = (int(card[0]) - 1) - flipped[0]
    scorer.set_dataset(flipped, card)
    assert np.allclose(scorer.score_adjacency(adj), base, rtol=1e-12, atol=1e-9)
    # duplicating the data doubles the log-likelihood
    scorer.set_dataset(codes, card)
    ll1 = scorer.score_adjacency(adj, metric="loglik")
    scorer.set_dataset(np.concatenate([codes, codes], axis=1), card)
    assert np.allclose(scorer.score_adjacency(adj, metric="loglik"), 2 * ll1, rtol=1e-12, atol=1e-9)
    # the empty graph is the sum of the marginal terms
    scorer.set_dataset(codes, card)
    empty = scorer.score_adjacency(np.zeros((1, n, n), dtype=np.uint8))[0]
    assert empty == pytest.approx(sum(O.family_score(codes, card, i, []) for i in range(n)), rel=1e-12, abs=1e-9)
