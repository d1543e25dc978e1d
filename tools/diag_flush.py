import sys, time, numpy as np, torch
sys.path.insert(0, '.')
dev = torch.device('cuda', 0)
def ev_time(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
b8 = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
b32 = torch.empty(64 << 20, dtype=torch.int32, device=dev)
print("zero_ uint8 256MB  : %.3f ms" % ev_time(lambda: b8.zero_()))
print("zero_ int32 256MB  : %.3f ms" % ev_time(lambda: b32.zero_()))
print("fill_ int32 256MB  : %.3f ms" % ev_time(lambda: b32.fill_(1)))
src = torch.empty(64 << 20, dtype=torch.int32, device=dev)
print("copy_ 256MB        : %.3f ms" % ev_time(lambda: b32.copy_(src)))
import bench, dags_vae_search_b200 as pkg
for w in ("asia", "sachs"):
    cfg = bench.WORKLOADS[w]
    _, card, codes = bench.make_dataset_gpu(cfg, cfg["rows"], dev)
    s = pkg.BicScorer(codes, card); s.set_stream(torch.cuda.current_stream().cuda_stream)
    adj = torch.from_numpy(bench.candidate_batch(cfg, cfg["batch"], 0, 0, 1)).to(dev)
    out = torch.empty(cfg["batch"], dtype=torch.float64, device=dev)
    def step():
        s.cache_clear(); s.score_adjacency_into(adj.data_ptr(), cfg["batch"], out.data_ptr(), device=True)
    def step_flush():
        b32.zero_(); step()
    print(w, "step: %.3f ms   flush+step: %.3f ms   cache bytes %d" % (ev_time(step), ev_time(step_flush), s.cache_stats()["bytes"]))
