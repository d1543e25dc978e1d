"""Worker for tests/test_gpu_multi.py: one process per GPU under torch.distributed.run.

Row sharding (BASELINE config 5): every rank holds a slice of the rows, partial count tables are
summed by ncclAllReduce(uint32) inside libbicgpu; counts must equal the oracle's on the full
dataset bit for bit, scores within 1e-9 relative, and all ranks must hold identical bits.
Candidate sharding (configs 1-4): dataset replicated, each rank scores its slice of the batch.
Family sharding: dataset replicated, every rank is given the same global batch and counts only the
families it owns; the all-reduced terms must reproduce a single-GPU run bit for bit.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dags_vae_search_b200 as pkg  # noqa: E402
from dags_vae_search_b200 import dist as bdist, synth  # noqa: E402
from oracle import c_oracle as C  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, N = 20, 400_003
    adj, card, cpts = synth.make_network(n, 30, 3, [2, 3, 4], seed=1)
    codes = synth.forward_sample(adj, card, cpts, N, np.random.default_rng(2))      # same on every rank
    dags = synth.er_candidates(n, 300, 19, 40, 5, seed=3)

    # ---- row sharding
    lo, hi = bdist.shard_range(N, rank, world)
    s = pkg.BicScorer(np.ascontiguousarray(codes[:, lo:hi]), card, device=local)
    bdist.init_row_sharding(s)
    s.profile_reset()
    got = s.score_adjacency(dags)
    fams = [(0, []), (3, [1]), (5, [0, 2, 7]), (9, [1, 2, 3, 4, 5, 6]), (11, [0, 1, 2, 3, 4, 5, 6, 7, 8])]
    tabs = s.count_families([f[0] for f in fams], [f[1] for f in fams])
    want = C.score_dags_adj(codes, card, dags)
    rel = np.abs(got - want) / np.abs(want)
    assert rel.max() < 1e-9, rel.max()
    for (i, ps), t in zip(fams, tabs):
        assert np.array_equal(t, C.family_counts(codes, card, i, ps)), (i, ps)
        assert t.sum() == N
    allbits = [torch.zeros(len(got), dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(allbits, torch.from_numpy(got).cuda())
    for other in allbits:
        assert torch.equal(other, allbits[0])          # every rank: identical bits
    # a second batch reuses cached families and still agrees
    dags2 = synth.er_candidates(n, 200, 19, 40, 5, seed=4)
    got2 = s.score_adjacency(np.concatenate([dags[:50], dags2]))
    want2 = np.concatenate([want[:50], C.score_dags_adj(codes, card, dags2)])
    assert (np.abs(got2 - want2) / np.abs(want2)).max() < 1e-9
    prof_rows = s.profile()
    s.end_row_sharding()
    s.close()
    # the same through ncclAllReduce(uint32) of the tables (BIC_NO_PUSH=1): identical bits
    os.environ["BIC_NO_PUSH"] = "1"
    s2 = pkg.BicScorer(np.ascontiguousarray(codes[:, lo:hi]), card, device=local)
    os.environ.pop("BIC_NO_PUSH")
    bdist.init_row_sharding(s2)
    assert np.array_equal(s2.score_adjacency(dags), got)
    s2.end_row_sharding()
    s2.close()
    # class 2 / class 3 tables and derived families through the exchange buffers (rows >= 2^20 per rank)
    N3 = 2_200_001
    card3 = np.array([21, 20, 19, 3, 3, 3, 2, 4], dtype=np.int32)
    rng3 = np.random.default_rng(6)
    codes3 = np.stack([rng3.integers(0, c, size=N3) for c in card3]).astype(np.uint8)
    lo3, hi3 = bdist.shard_range(N3, rank, world)
    s3 = pkg.BicScorer(np.ascontiguousarray(codes3[:, lo3:hi3]), card3, device=local)
    bdist.init_row_sharding(s3)
    fams3 = [(3, [0, 1]), (4, [0, 1, 2]), (3, [0, 1, 2, 5]), (6, [7]), (7, [3, 4, 6]), (6, [3, 7]), (5, [])]
    node3 = np.array([f[0] for f in fams3], dtype=np.int32)
    off3 = np.zeros(len(fams3) + 1, dtype=np.int64)
    off3[1:] = np.cumsum([len(f[1]) for f in fams3])
    par3 = np.array([p for f in fams3 for p in f[1]], dtype=np.int32)
    s3.profile_reset()
    got3 = s3.score_families([f[0] for f in fams3], [f[1] for f in fams3])
    want3 = C.score_families(codes3, card3, node3, off3, par3)
    assert (np.abs(got3 - want3) / np.abs(want3)).max() < 1e-9
    assert s3.profile()["families_derived"] > 0
    for (i, ps), t in zip(fams3[:3], s3.count_families([f[0] for f in fams3[:3]], [f[1] for f in fams3[:3]])):
        assert np.array_equal(t, C.family_counts(codes3, card3, i, ps)) and t.sum() == N3
    s3.end_row_sharding()
    s3.close()

    # ---- candidate sharding
    full = pkg.BicScorer(codes, card, device=local)
    blo, bhi = bdist.shard_range(len(dags), rank, world)
    mine = full.score_adjacency(dags[blo:bhi])
    gathered = bdist.gather_scores(mine, len(dags))
    assert (np.abs(gathered - want) / np.abs(want)).max() < 1e-9
    full.close()

    # ---- family sharding over a global batch (dataset replicated, derive path on: N >= 2^20)
    N2 = 1_100_000
    codes2 = synth.forward_sample(adj, card, cpts, N2, np.random.default_rng(5))
    single = pkg.BicScorer(codes2, card, device=local)
    want_bits = single.score_adjacency(dags)
    single.close()
    fam = pkg.BicScorer(codes2, card, device=local)
    bdist.init_family_sharding(fam)
    mine_cuda = torch.from_numpy(dags[blo:bhi] if (bhi - blo) * world == len(dags) else dags[:len(dags) // world]).cuda()
    if (bhi - blo) * world == len(dags):
        glob = bdist.all_gather_batches(mine_cuda)
        assert np.array_equal(glob.cpu().numpy(), dags)
    fam.profile_reset()
    got_f = fam.score_adjacency(dags)
    assert np.array_equal(got_f, want_bits)                 # bit-identical to one GPU
    prof = fam.profile()
    counted = torch.tensor([prof["families_counted"], prof["families_derived"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(counted)
    st = fam.cache_stats()
    assert int(counted.sum()) == st["families"] and prof["families_counted"] < st["families"]   # the work was split
    if (bhi - blo) * world == len(dags):                    # every rank passes only its own DAGs (BIC_FLAG_LOCAL_BATCH)
        fam.cache_clear()
        mine_scores = fam.score_adjacency_local(mine_cuda)
        assert np.array_equal(mine_scores.cpu().numpy(), want_bits[blo:bhi])
    got_f2 = fam.score_adjacency(dags2)                     # second batch: cached + new families
    one = pkg.BicScorer(codes2, card, device=local)
    one.score_adjacency(dags)
    assert np.array_equal(got_f2, one.score_adjacency(dags2))
    one.close()
    fam.close()
    dist.barrier()
    if rank == 0:
        print("row-sharded exchange steps: fused reduce-scatter %d, ncclAllReduce %d" % (prof_rows["exchange_fused"], prof_rows["exchange_nccl"]))
        print("multigpu ok", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
