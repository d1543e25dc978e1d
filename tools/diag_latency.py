"""Per-call latency of the drop-in BNLearnWrapper.score() (one DAG per call, like the reference)."""
import sys, time, numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import dags_vae_search_b200 as pkg
from dags_vae_search_b200 import synth
from graph_stub import Graph
ev = pkg.BNLearnWrapper("asia", "bic")
adjs = synth.er_candidates(8, 2000, 5, 14, None, seed=1)
graphs = [Graph(8, list(zip(*np.nonzero(a))), list(range(8))) for a in adjs]
for g in graphs[:200]: ev.score(g)
t0 = time.perf_counter()
for g in graphs: ev.score(g)
dt = time.perf_counter() - t0
print("BNLearnWrapper.score: %.1f us per DAG (%d calls, warm cache)" % (dt / len(graphs) * 1e6, len(graphs)))
s = ev.scorer
one = adjs[:1].copy()
t0 = time.perf_counter()
for i in range(2000): s.score_adjacency(adjs[i:i+1])
print("scorer.score_adjacency(B=1): %.1f us per call" % ((time.perf_counter() - t0) / 2000 * 1e6))
t0 = time.perf_counter()
for i in range(200): s.score_adjacency(adjs[:150])
print("scorer.score_adjacency(B=150): %.1f us per call" % ((time.perf_counter() - t0) / 200 * 1e6))
print(s.profile()["kernel_launches"])
