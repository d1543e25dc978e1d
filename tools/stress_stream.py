"""Long warm stream on the bench workload: cache growth through rehashes, stable memory, and
scores of the first batch reproduced bit for bit from the cache at the end."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench, dags_vae_search_b200 as pkg
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
cfg = bench.WORKLOADS['alarm']
dev = torch.device('cuda', 0)
_, card, codes = bench.make_dataset_gpu(cfg, cfg['rows'], dev)
s = pkg.BicScorer(codes, card)
del codes
first = None
t0 = time.perf_counter()
free0 = torch.cuda.mem_get_info()[0]
for i in range(steps):
    adj = bench.candidate_batch(cfg, 4096, i, 0, 1)
    out = s.score_adjacency(adj)
    assert not np.isnan(out).any()
    if i == 0:
        first, first_adj = out.copy(), adj
    if i % 20 == 19:
        st = s.cache_stats()
        print(f"step {i+1}: {(i+1)*4096/(time.perf_counter()-t0):.0f} DAGs/s incl. host candidate generation, cache {st['families']} families, {st['bytes']/1e6:.0f} MB, free HBM {torch.cuda.mem_get_info()[0]/1e9:.1f} GB", flush=True)
again = s.score_adjacency(first_adj)
assert np.array_equal(again, first)
st = s.cache_stats()
assert st["misses"] == st["families"]
print("ok", st)
