import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench, dags_vae_search_b200 as pkg
cfg = bench.WORKLOADS['alarm']
dev = torch.device('cuda', 0)
_, card, codes = bench.make_dataset_gpu(cfg, cfg['rows'], dev)
s = pkg.BicScorer(codes, card)
s.profile_enable(True)
batches = [torch.from_numpy(bench.candidate_batch(cfg, 4096, i, 0, 1)).to(dev) for i in range(6)]
out = torch.empty(4096, dtype=torch.float64, device=dev)
for mode in ('cold', 'warm'):
    s.cache_clear()
    for i, b in enumerate(batches):
        if mode == 'cold': s.cache_clear()
        s.profile_reset()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s.score_adjacency_into(b.data_ptr(), 4096, out.data_ptr(), device=True)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
        p = s.profile(); st = s.cache_stats()
        print(mode, i, f"{dt:7.1f} ms count_ms {p['count_ms']:7.1f} counted {p['families_counted']:6d} derived {p['families_derived']:6d} class_fam {p['class_families']} class_ms {[round(x,1) for x in p['class_ms']]} cache {st['families']}")
