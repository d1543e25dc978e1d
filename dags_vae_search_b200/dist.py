"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` for rendezvous only.

* candidate sharding (BASELINE configs 1-4): the dataset is replicated, each rank scores its
  slice of the candidate batch, no data-path collective; ``shard_range`` gives the slice and
  ``gather_scores`` concatenates the fp64 results when a caller needs them in one place.
* row sharding (config 5): each rank holds N/world rows; ``init_row_sharding`` creates the
  NCCL communicator inside libbicgpu (id from rank 0, broadcast through the process group)
  so that partial count tables are summed with ncclAllReduce(uint32) over NVLink.
"""
from __future__ import annotations

import ctypes
from typing import Tuple

import numpy as np

from . import _native as nat


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of ``total`` items for ``rank``."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scores(local: np.ndarray, total: int, group=None) -> np.ndarray:
    """All ranks' score slices (``shard_range`` order) -> the full float64 [total] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    buf = torch.zeros(width, dtype=torch.float64, device=dev)
    lo, hi = sizes[rank]
    buf[:hi - lo] = torch.as_tensor(np.asarray(local, dtype=np.float64), device=dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return np.concatenate([parts[r][:sizes[r][1] - sizes[r][0]].cpu().numpy() for r in range(world)])


def broadcast_unique_id(group=None) -> bytes:
    """Rank 0 makes the NCCL unique id inside libbicgpu; everyone receives its 128 bytes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    buf = (ctypes.c_uint8 * 128)()
    if rank == 0:
        rc = nat.lib().bic_comm_unique_id(ctypes.addressof(buf))
        if rc != nat.BIC_OK:
            raise nat.BicError(rc, (nat.lib().bic_last_error(None) or b"").decode())
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().tolist())


def init_row_sharding(scorer, group=None) -> None:
    import torch.distributed as dist
    uid = broadcast_unique_id(group)
    scorer.init_row_sharding(dist.get_rank(group), dist.get_world_size(group), uid)


def init_family_sharding(scorer, group=None) -> None:
    """Families of a global candidate batch are split over the ranks (dataset replicated); call the
    scorer with the SAME batch on every rank, e.g. after ``all_gather_batches``."""
    import torch.distributed as dist
    uid = broadcast_unique_id(group)
    scorer.init_family_sharding(dist.get_rank(group), dist.get_world_size(group), uid)


def all_gather_batches(local_adj, group=None):
    """Concatenate every rank's CUDA adjacency batch ``[B, n, n]`` (same B everywhere) in rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * local_adj.shape[0],) + tuple(local_adj.shape[1:]), dtype=local_adj.dtype, device=local_adj.device)
    dist.all_gather_into_tensor(out, local_adj.contiguous(), group=group)
    return out
