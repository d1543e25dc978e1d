import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench, dags_vae_search_b200 as pkg
from dags_vae_search_b200 import dist as bdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1]          # "default" (torch current stream) | "own" | "side" (a torch side stream)
def log(*a):
    print(f"[r{rank} {mode}]", *a, flush=True)
cfg = bench.WORKLOADS["alarm"]
_, card, codes = bench.make_dataset_gpu(cfg, 1_200_000, dev)
s = pkg.BicScorer(codes, card, device=local)
side = torch.cuda.Stream()
if mode == "default": s.set_stream(torch.cuda.current_stream().cuda_stream)
if mode == "side": s.set_stream(side.cuda_stream)
bdist.init_family_sharding(s)
log("init done")
B = 512
adjs = [torch.from_numpy(bench.candidate_batch(cfg, B, i, rank, world)).to(dev) for i in range(6)]
out = torch.empty(B * world, dtype=torch.float64, device=dev)
for i, a in enumerate(adjs):
    s.cache_clear()
    if mode == "side":
        with torch.cuda.stream(side):
            glob = bdist.all_gather_batches(a)
            s.score_adjacency_into(glob.data_ptr(), B * world, out.data_ptr(), device=True)
    else:
        glob = bdist.all_gather_batches(a)
        if mode == "own": torch.cuda.current_stream().synchronize()
        s.score_adjacency_into(glob.data_ptr(), B * world, out.data_ptr(), device=True)
    log("step", i, float(out.sum().item()))
dist.barrier(); torch.cuda.synchronize()
log("done")
dist.destroy_process_group()
