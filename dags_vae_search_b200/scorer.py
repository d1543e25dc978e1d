"""``BicScorer`` — batched host API over the C ABI (``include/bicgpu.h``).

One scorer = one GPU context holding the dataset in HBM plus the family-score cache.  Inputs are
numpy arrays (host pointers; the library stages them) or CUDA ``torch.Tensor`` objects (device
pointers, tensors are only buffer carriers).  There is no CPU path.

Reference behaviour covered: ``BNLearnWrapper.score`` (``src/problem/bn/bnlearn.py:27-61``) and
the R child it spawns (``bnlearn_score.R:7-40``), generalised from one DAG per process to a
batch per call.
"""
from __future__ import annotations

import ctypes
from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as nat


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _current_stream(t) -> int:
    """The CUDA stream torch would launch on for tensor ``t``'s device (the producer of ``t``)."""
    import torch
    return int(torch.cuda.current_stream(t.device).cuda_stream)


def _family_csr(nodes: Sequence[int], parent_lists: Sequence[Iterable[int]]):
    node = np.ascontiguousarray(nodes, dtype=np.int32)
    lens = np.fromiter((len(p) for p in parent_lists), dtype=np.int64, count=len(parent_lists))
    off = np.zeros(len(node) + 1, dtype=np.int64)
    np.cumsum(lens, out=off[1:])
    flat = np.fromiter((int(x) for p in parent_lists for x in p), dtype=np.int32, count=int(off[-1]))
    if flat.size == 0:
        flat = np.zeros(1, dtype=np.int32)
    return node, off, flat


class BicScorer:
    """Decomposable discrete BN scores (bic / aic / loglik / bde / k2) of DAG batches on one B200."""

    def __init__(self, codes, card, device: int = 0, metric: str = "bic", iss: float = 1.0):
        if metric not in nat.METRICS:
            raise NotImplementedError(f"metric {metric!r}: only {sorted(nat.METRICS)} are implemented")
        self._lib = nat.lib()
        self._ctx = ctypes.c_void_p()
        self.metric = metric
        self.device = int(device)
        rc = self._lib.bic_create(ctypes.byref(self._ctx), self.device)
        if rc != nat.BIC_OK:
            msg = self._lib.bic_last_error(None)
            self._ctx = ctypes.c_void_p()
            raise nat.BicError(rc, msg.decode() if msg else "")
        self._iss = 1.0
        self._shard = (0, 1)     # (rank, world) when the rows are sharded
        self.set_dataset(codes, card)
        if iss != 1.0:
            self.set_iss(iss)

    def set_iss(self, iss: float) -> None:
        """Imaginary sample size of the ``bde`` (BDeu) metric, bnlearn's ``iss`` (default 1)."""
        self._check(self._lib.bic_set_iss(self._ctx, float(iss)))
        self._iss = float(iss)

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            self._lib.bic_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        nat.check(self._ctx, rc)

    def _after_producer(self, tensor) -> None:
        """Device inputs: whatever torch queued on its current stream (the decoder, a dtype
        conversion, an all-gather) must finish before the library's kernels read the buffer.  The
        library runs on its own stream, so the ordering is made explicit: an event on torch's
        stream that the context's stream waits on (``bic_wait_stream``; no host synchronisation)."""
        self._check(self._lib.bic_wait_stream(self._ctx, ctypes.c_void_p(_current_stream(tensor))))

    # ------------------------------------------------------------------- dataset
    def set_dataset(self, codes, card) -> None:
        """codes: uint8 ``[n, N]`` (variable-major, i.e. column-major samples); card: ``[n]``."""
        card = np.ascontiguousarray(card, dtype=np.int32)
        if _is_torch(codes):
            if codes.dtype.__str__() != "torch.uint8" or codes.dim() != 2:
                raise ValueError("codes tensor must be uint8 [n, N]")
            if not codes.is_cuda:
                codes = codes.numpy()
        if _is_torch(codes):
            codes = codes.contiguous()
            n, N = int(codes.shape[0]), int(codes.shape[1])
            ptr, is_dev, stride = codes.data_ptr(), 1, N
            self._after_producer(codes)
        else:
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            if codes.ndim != 2:
                raise ValueError("codes must be [n, N]")
            n, N = codes.shape
            ptr, is_dev, stride = codes.ctypes.data, 0, N
        if len(card) != n:
            raise ValueError(f"card has {len(card)} entries for {n} variables")
        self._check(self._lib.bic_set_dataset(self._ctx, ptr, N, n, stride, card.ctypes.data, is_dev))
        self.n, self.N = int(n), int(N)
        self.card = card.copy()

    # ------------------------------------------------------------------ families
    def family_cells(self, node: int, parents: Iterable[int]) -> Tuple[int, int]:
        q = 1
        for p in set(int(x) for x in parents):
            q *= int(self.card[p])
        return q, int(self.card[int(node)])

    def count_families(self, nodes: Sequence[int], parent_lists: Sequence[Iterable[int]]) -> List[np.ndarray]:
        """Dense int32 contingency tables ``[q, r]`` (cell = j * r + x), straight from the count
        kernel, bypassing the cache."""
        node, off, flat = _family_csr(nodes, parent_lists)
        shapes = [self.family_cells(i, p) for i, p in zip(nodes, parent_lists)]
        coff = np.zeros(len(node) + 1, dtype=np.int64)
        np.cumsum([q * r for q, r in shapes], out=coff[1:])
        out = np.zeros(max(int(coff[-1]), 1), dtype=np.int32)
        self._check(self._lib.bic_count_families(self._ctx, node.ctypes.data, off.ctypes.data, flat.ctypes.data,
                                                 len(node), coff.ctypes.data, out.ctypes.data, 0))
        return [out[coff[f]:coff[f + 1]].reshape(shapes[f]) for f in range(len(node))]

    def family_counts(self, node: int, parents: Iterable[int]) -> np.ndarray:
        return self.count_families([node], [list(parents)])[0]

    def score_families(self, nodes: Sequence[int], parent_lists: Sequence[Iterable[int]],
                       metric: Optional[str] = None, no_cache: bool = False) -> np.ndarray:
        node, off, flat = _family_csr(nodes, parent_lists)
        out = np.zeros(len(node), dtype=np.float64)
        flags = (nat.FLAG_NO_CACHE if no_cache else 0) | (0 if self.derive else nat.FLAG_NO_DERIVE)
        self._check(self._lib.bic_score_families(self._ctx, node.ctypes.data, off.ctypes.data, flat.ctypes.data,
                                                 len(node), self._metric(metric), out.ctypes.data, flags))
        return out

    def score_families_csr(self, node: np.ndarray, off: np.ndarray, parents: np.ndarray,
                           metric: Optional[str] = None, no_cache: bool = False) -> np.ndarray:
        node = np.ascontiguousarray(node, dtype=np.int32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        parents = np.ascontiguousarray(parents, dtype=np.int32)
        if parents.size == 0:
            parents = np.zeros(1, dtype=np.int32)
        out = np.zeros(len(node), dtype=np.float64)
        flags = (nat.FLAG_NO_CACHE if no_cache else 0) | (0 if self.derive else nat.FLAG_NO_DERIVE)
        self._check(self._lib.bic_score_families(self._ctx, node.ctypes.data, off.ctypes.data, parents.ctypes.data,
                                                 len(node), self._metric(metric), out.ctypes.data, flags))
        return out

    # ---------------------------------------------------------------------- DAGs
    def _metric(self, metric: Optional[str]) -> int:
        m = self.metric if metric is None else metric
        if m not in nat.METRICS:
            raise NotImplementedError(f"metric {m!r}: only {sorted(nat.METRICS)} are implemented")
        return nat.METRICS[m]

    #: set to False to count every family from the rows (no marginalisation from supersets)
    derive = True

    def _flags(self, check_acyclic: bool, no_cache: bool, device: bool) -> int:
        return ((0 if check_acyclic else nat.FLAG_NO_CYCLE_CHECK) | (nat.FLAG_NO_CACHE if no_cache else 0)
                | (nat.FLAG_DEVICE_PTRS if device else 0) | (0 if self.derive else nat.FLAG_NO_DERIVE))

    def score_adjacency(self, adj, metric: Optional[str] = None, check_acyclic: bool = True,
                        no_cache: bool = False, return_invalid: bool = False):
        """adj: uint8 ``[B, n, n]`` (or ``[n, n]``), ``adj[b, p, c] != 0`` <=> edge p -> c
        (row = parent, as ``bnlearn.py:44`` serialises it).  Returns float64 ``[B]``; cyclic DAGs
        score NaN.  CUDA tensors in -> CUDA tensor out."""
        n = self.n
        inv = ctypes.c_int64(0)
        if _is_torch(adj) and adj.is_cuda:
            import torch
            a = adj.reshape(-1, n, n).to(torch.uint8).contiguous()
            out = torch.empty(a.shape[0], dtype=torch.float64, device=a.device)
            self._after_producer(a)
            self._check(self._lib.bic_score_dags_adj(self._ctx, a.data_ptr(), a.shape[0], self._metric(metric),
                                                     out.data_ptr(), ctypes.byref(inv),
                                                     self._flags(check_acyclic, no_cache, True)))
        else:
            if _is_torch(adj):
                adj = adj.numpy()
            a = np.ascontiguousarray(adj, dtype=np.uint8).reshape(-1, n, n)
            out = np.empty(a.shape[0], dtype=np.float64)
            self._check(self._lib.bic_score_dags_adj(self._ctx, a.ctypes.data, a.shape[0], self._metric(metric),
                                                     out.ctypes.data, ctypes.byref(inv),
                                                     self._flags(check_acyclic, no_cache, False)))
        return (out, int(inv.value)) if return_invalid else out

    def score_adjacency_into(self, adj_ptr: int, B: int, out_ptr: int, device: bool, metric: Optional[str] = None,
                             check_acyclic: bool = True, no_cache: bool = False, extra_flags: int = 0) -> int:
        """Raw-pointer variant (what bench.py times): no allocation, no conversion, no stream ordering —
        device inputs must be complete (see ``bic_wait_stream``).  ``extra_flags``: e.g.
        ``FLAG_LOCAL_BATCH`` for a family-sharded scorer."""
        inv = ctypes.c_int64(0)
        self._check(self._lib.bic_score_dags_adj(self._ctx, adj_ptr, B, self._metric(metric), out_ptr, ctypes.byref(inv),
                                                 self._flags(check_acyclic, no_cache, device) | int(extra_flags)))
        return int(inv.value)

    def score_csr_into(self, off_ptr: int, parents_ptr: int, B: int, out_ptr: int, device: bool,
                       metric: Optional[str] = None, check_acyclic: bool = True, no_cache: bool = False) -> int:
        """Raw-pointer CSR variant (int64 offsets [B*n+1], int32 parents)."""
        inv = ctypes.c_int64(0)
        self._check(self._lib.bic_score_dags_csr(self._ctx, off_ptr, parents_ptr, B, self._metric(metric), out_ptr,
                                                 ctypes.byref(inv), self._flags(check_acyclic, no_cache, device)))
        return int(inv.value)

    def score_csr(self, off, parents, B: int, metric: Optional[str] = None, check_acyclic: bool = True,
                  no_cache: bool = False, return_invalid: bool = False):
        """Parent lists in CSR: family (b, i) = parents[off[b*n+i] : off[b*n+i+1]]."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        parents = np.ascontiguousarray(parents, dtype=np.int32)
        if off.shape[0] != B * self.n + 1:
            raise ValueError("off must have B*n+1 entries")
        if parents.size == 0:
            parents = np.zeros(1, dtype=np.int32)
        out = np.empty(B, dtype=np.float64)
        inv = ctypes.c_int64(0)
        self._check(self._lib.bic_score_dags_csr(self._ctx, off.ctypes.data, parents.ctypes.data, B,
                                                 self._metric(metric), out.ctypes.data, ctypes.byref(inv),
                                                 self._flags(check_acyclic, no_cache, False)))
        return (out, int(inv.value)) if return_invalid else out

    def score_wire(self, labels, ebits, metric: Optional[str] = None, no_cache: bool = False,
                   return_invalid: bool = False):
        """Reference candidate wire format (``src/toolkit/labeled.py:116-154``): ``labels[b, v]`` =
        BN variable of vertex v (uint16, any n), ``ebits[b, v, w]`` = 32-bit edge words, bit ``u % 32``
        of word ``u // 32`` <=> edge vertex u -> vertex v (``[B, n]`` is accepted for n <= 32).
        CUDA tensors in -> CUDA tensor out (decoder output that never leaves the GPU)."""
        inv = ctypes.c_int64(0)
        n = self.n
        ew_min = (n + 31) // 32
        if _is_torch(labels) and labels.is_cuda:
            import torch
            lab = labels.reshape(-1, n).to(torch.int64)
            B = lab.shape[0]
            # a label outside 0..65535 must not wrap into the valid range: 65535 is rejected by the kernel (n <= 1024)
            lab = torch.where((lab < 0) | (lab > 65535), torch.full_like(lab, 65535), lab)
            lab16 = lab.to(torch.uint16).contiguous()
            eb = ebits.reshape(B, n, -1).to(torch.int32).contiguous()       # bit pattern of the uint32 words
            if eb.shape[2] < ew_min:
                raise ValueError(f"ebits needs {ew_min} words per vertex for n = {n}")
            out = torch.empty(B, dtype=torch.float64, device=lab16.device)
            self._after_producer(eb)
            self._check(self._lib.bic_score_dags_wire16(self._ctx, lab16.data_ptr(), eb.data_ptr(), eb.shape[2], B,
                                                        self._metric(metric), out.data_ptr(), ctypes.byref(inv),
                                                        self._flags(True, no_cache, True)))
            return (out, int(inv.value)) if return_invalid else out
        labels = np.asarray(labels)
        if labels.size and (labels.min() < 0 or labels.max() > 65535):
            labels = np.where((labels < 0) | (labels > 65535), 65535, labels)
        labels = np.ascontiguousarray(labels, dtype=np.uint16).reshape(-1, n)
        ebits = np.ascontiguousarray(ebits, dtype=np.uint32).reshape(labels.shape[0], n, -1)
        if ebits.shape[2] < ew_min:
            raise ValueError(f"ebits needs {ew_min} words per vertex for n = {n}")
        out = np.empty(labels.shape[0], dtype=np.float64)
        self._check(self._lib.bic_score_dags_wire16(self._ctx, labels.ctypes.data, ebits.ctypes.data, ebits.shape[2],
                                                    labels.shape[0], self._metric(metric), out.ctypes.data,
                                                    ctypes.byref(inv), self._flags(True, no_cache, False)))
        return (out, int(inv.value)) if return_invalid else out

    def score_adjacency_local(self, adj, metric: Optional[str] = None, check_acyclic: bool = True):
        """Family-sharded scoring of a sharded candidate batch: ``adj`` holds THIS rank's B DAGs (CUDA
        tensor ``[B, n, n]``, same B on every rank).  The family keys are all-gathered over NVLink
        inside the library, the union is deduplicated identically on every rank, each rank counts only
        the families it owns and returns the scores of its own B DAGs."""
        import torch
        n = self.n
        a = adj.reshape(-1, n, n).to(torch.uint8).contiguous()
        out = torch.empty(a.shape[0], dtype=torch.float64, device=a.device)
        inv = ctypes.c_int64(0)
        self._after_producer(a)
        self._check(self._lib.bic_score_dags_adj(self._ctx, a.data_ptr(), a.shape[0], self._metric(metric), out.data_ptr(),
                                                 ctypes.byref(inv),
                                                 self._flags(check_acyclic, False, True) | nat.FLAG_LOCAL_BATCH))
        return out

    # --------------------------------------------------------------------- cache
    def cache_clear(self) -> None:
        self._check(self._lib.bic_cache_clear(self._ctx))

    def cache_reserve(self, families: int) -> None:
        self._check(self._lib.bic_cache_reserve(self._ctx, int(families)))

    def cache_stats(self) -> dict:
        s = nat.CacheStats()
        self._check(self._lib.bic_cache_stats(self._ctx, ctypes.byref(s)))
        return {k: int(getattr(s, k)) for k, _ in nat.CacheStats._fields_}

    def save_cache(self, path: str) -> int:
        """Checkpoint the family-score cache (keys + terms) to an ``.npz``; returns the family count."""
        fam, kind = ctypes.c_int64(0), ctypes.c_int32(0)
        self._check(self._lib.bic_cache_export(self._ctx, None, None, None, 0, ctypes.byref(fam), ctypes.byref(kind)))
        F, Wk = int(fam.value), 1 + (self.n + 63) // 64
        keys = np.zeros((max(F, 1), Wk), dtype=np.uint64)
        terms = np.zeros(max(F, 1), dtype=np.float64)
        nparams = np.zeros(max(F, 1), dtype=np.float64)
        if F:
            self._check(self._lib.bic_cache_export(self._ctx, keys.ctypes.data, terms.ctypes.data, nparams.ctypes.data, F,
                                                   ctypes.byref(fam), ctypes.byref(kind)))
        np.savez_compressed(path, keys=keys[:F], terms=terms[:F], nparams=nparams[:F], kind=np.int32(kind.value),
                            dataset_tag=np.array(self._dataset_tag(int(kind.value))), n=np.int64(self.n), N=np.int64(self.N))
        return F

    def _dataset_tag(self, kind: int) -> str:
        """What a cache checkpoint is valid for: the dataset's content (device-side fingerprint of every
        code), its shape and cardinalities, the sharding of the rows, and ``iss`` for bde terms."""
        import hashlib
        fp = ctypes.c_uint64(0)
        self._check(self._lib.bic_dataset_fingerprint(self._ctx, ctypes.byref(fp)))
        h = hashlib.sha256(np.asarray([self.n, self.N, self._shard[0], self._shard[1]], dtype=np.int64).tobytes())
        h.update(np.asarray([fp.value], dtype=np.uint64).tobytes())
        h.update(self.card.tobytes())
        if kind == 1:
            h.update(np.asarray([self._iss], dtype=np.float64).tobytes())
        return h.hexdigest()

    def load_cache(self, path: str) -> int:
        """Resume from ``save_cache``: replaces the cache content.  Refuses a checkpoint made with
        another dataset shape / cardinalities."""
        d = np.load(path)
        if str(d["dataset_tag"]) != self._dataset_tag(int(d["kind"])):
            raise ValueError("cache checkpoint was made with a different dataset (content, shape, cardinalities, "
                             "row sharding or iss differ)")
        keys = np.ascontiguousarray(d["keys"], dtype=np.uint64)
        terms = np.ascontiguousarray(d["terms"], dtype=np.float64)
        nparams = np.ascontiguousarray(d["nparams"], dtype=np.float64)
        self._check(self._lib.bic_cache_import(self._ctx, keys.ctypes.data, terms.ctypes.data, nparams.ctypes.data,
                                               len(terms), int(d["kind"])))
        return len(terms)

    # ----------------------------------------------------------------- profiling
    def profile_enable(self, on: bool = True) -> None:
        self._check(self._lib.bic_profile_enable(self._ctx, 1 if on else 0))

    def profile_reset(self) -> None:
        self._check(self._lib.bic_profile_reset(self._ctx))

    def profile(self) -> dict:
        p = nat.Profile()
        self._check(self._lib.bic_profile_get(self._ctx, ctypes.byref(p)))
        out = {}
        for k, _ in nat.Profile._fields_:
            v = getattr(p, k)
            out[k] = list(v) if hasattr(v, "__len__") else (float(v) if k.endswith("_ms") else int(v))
        return out

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        self._check(self._lib.bic_set_stream(self._ctx, ctypes.c_void_p(cuda_stream or 0)))

    def sync(self) -> None:
        self._check(self._lib.bic_sync(self._ctx))

    # -------------------------------------------------------------- row sharding
    def init_row_sharding(self, rank: int, world: int, unique_id: bytes) -> None:
        """Rows are sharded over ``world`` ranks (one process per GPU): count tables are summed
        with an NCCL uint32 all-reduce before the fp64 reduce.  See ``dist.py``."""
        buf = (ctypes.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self._lib.bic_comm_init(self._ctx, ctypes.addressof(buf), int(rank), int(world)))
        self._shard = (int(rank), int(world))

    def init_family_sharding(self, rank: int, world: int, unique_id: bytes) -> None:
        """Dataset replicated, every rank is given the same global candidate batch; each rank counts
        only the families it owns and the terms are all-reduced (bit-identical to one GPU)."""
        self.init_row_sharding(rank, world, unique_id)
        self._check(self._lib.bic_comm_mode(self._ctx, 1))
        self._shard = (0, 1)     # every rank holds all rows

    def end_row_sharding(self) -> None:
        self._check(self._lib.bic_comm_destroy(self._ctx))
        self._shard = (0, 1)
