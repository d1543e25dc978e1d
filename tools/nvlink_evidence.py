#!/usr/bin/env python
"""NVLink bytes of the row-sharded exchange, read from the hardware counters.

Run under torch.distributed.run on N >= 2 GPUs of one box.  Rank 0 reads `nvidia-smi nvlink -gt d`
(cumulative data Tx / Rx KiB per link) on its GPU before and after a block of row-sharded steps,
once with the fused reduce-scatter (count kernels store partial tables into the owner rank's
exchange buffer) and once with ncclAllReduce(uint32) of the tables (BIC_NO_PUSH=1), and prints the
measured bytes per step next to what the library says it sent (bic_profile_t.exchange_bytes).
ncu / nsys cannot be wrapped around a multi-rank run on this pool; the link counters can be read.
"""
import json
import os
import re
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dags_vae_search_b200 as pkg  # noqa: E402
from dags_vae_search_b200 import dist as bdist  # noqa: E402


def link_kib(index):
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True).stdout
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return tx, rx, len(re.findall(r"Data Tx:", out))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    cfg = bench.WORKLOADS["diabetes"]
    rows = int(os.environ.get("EVIDENCE_ROWS", cfg["rows"]))
    steps = int(os.environ.get("EVIDENCE_STEPS", "200"))
    _, card, codes = bench.make_dataset_gpu(cfg, rows, device, shard=rank)
    n, batch = cfg["n"], cfg["batch"]
    adj = bench.candidate_batch(cfg, batch, 0, 0, 1)
    b, p, c = np.nonzero(adj.transpose(0, 2, 1))
    counts = np.bincount(b * n + p, minlength=batch * n)
    off = np.zeros(batch * n + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    d_off, d_par = torch.from_numpy(off).to(device), torch.from_numpy(c.astype(np.int32)).to(device)
    out = torch.empty(batch, dtype=torch.float64, device=device)
    result = {"world": world, "rows_per_gpu": rows, "steps": steps, "dags_per_step": batch}
    for mode in ("fused", "nccl"):
        if mode == "nccl":
            os.environ["BIC_NO_PUSH"] = "1"
        s = pkg.BicScorer(codes, card, device=local)
        os.environ.pop("BIC_NO_PUSH", None)
        bdist.init_row_sharding(s)
        for _ in range(3):
            s.cache_clear()
            s.score_csr_into(d_off.data_ptr(), d_par.data_ptr(), batch, out.data_ptr(), device=True)
        s.profile_enable(True)
        s.profile_reset()
        dist.barrier()
        torch.cuda.synchronize()
        before = link_kib(local) if rank == 0 else None
        dist.barrier()
        for _ in range(steps):
            s.cache_clear()
            s.score_csr_into(d_off.data_ptr(), d_par.data_ptr(), batch, out.data_ptr(), device=True)
        dist.barrier()
        torch.cuda.synchronize()
        after = link_kib(local) if rank == 0 else None
        prof = s.profile()
        if rank == 0:
            result[mode] = {"nvlink_tx_bytes_per_step": (after[0] - before[0]) * 1024 / steps,
                            "nvlink_rx_bytes_per_step": (after[1] - before[1]) * 1024 / steps, "links_read": after[2],
                            "library_exchange_bytes_per_step": prof["exchange_bytes"] / steps,
                            "exchange_steps_fused": prof["exchange_fused"], "exchange_steps_nccl": prof["exchange_nccl"],
                            "exchange_ms_per_step": prof["exchange_ms"] / steps, "count_ms_per_step": prof["count_ms"] / steps}
        s.end_row_sharding()
        s.close()
    if rank == 0:
        print(json.dumps(result))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
