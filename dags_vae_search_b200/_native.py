"""ctypes binding of ``csrc/libbicgpu.so`` (C ABI: ``include/bicgpu.h``).

The product path has no CPU fallback: if the CUDA library is missing or no B200 is visible,
loading / context creation raises.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.environ.get("BIC_LIB") or os.path.join(CSRC, "libbicgpu.so")   # BIC_LIB: an experimental build of the same ABI

BIC_OK = 0
FLAG_DEVICE_PTRS = 1
FLAG_NO_CYCLE_CHECK = 2
FLAG_NO_CACHE = 4
FLAG_NO_DERIVE = 8
FLAG_LOCAL_BATCH = 16
METRICS = {"bic": 0, "loglik": 1, "aic": 2, "bde": 3, "k2": 4}

STATUS_NAMES = {
    0: "BIC_OK", -1: "BIC_ERR_CUDA", -2: "BIC_ERR_ARG", -3: "BIC_ERR_NO_DATASET",
    -4: "BIC_ERR_TABLE_TOO_LARGE", -5: "BIC_ERR_OOM", -6: "BIC_ERR_NCCL", -7: "BIC_ERR_BAD_CODE",
    -8: "BIC_ERR_BAD_FAMILY",
}

# every symbol include/bicgpu.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "bic_version", "bic_build_info", "bic_create", "bic_destroy", "bic_last_error", "bic_set_stream", "bic_wait_stream", "bic_sync", "bic_set_iss",
    "bic_dataset_fingerprint", "bic_score_dags_wire16",
    "bic_set_dataset", "bic_count_families", "bic_score_families", "bic_score_dags_adj",
    "bic_score_dags_csr", "bic_score_dags_wire", "bic_cache_clear", "bic_cache_reserve",
    "bic_cache_stats", "bic_cache_export", "bic_cache_import", "bic_profile_enable", "bic_profile_reset", "bic_profile_get",
    "bic_comm_unique_id", "bic_comm_init", "bic_comm_destroy", "bic_comm_mode", "bic_plan_slices", "bic_range_plan",
]


class BicError(Exception):
    """Backend failure (the reference raises a bare ``Exception`` when its R child fails,
    ``src/problem/bn/bnlearn.py:56-57``)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class CacheStats(ctypes.Structure):
    _fields_ = [("families", ctypes.c_int64), ("capacity", ctypes.c_int64), ("lookups", ctypes.c_int64),
                ("misses", ctypes.c_int64), ("bytes", ctypes.c_int64)]


class PlanIn(ctypes.Structure):
    _fields_ = [("sm_count", ctypes.c_int32), ("N", ctypes.c_int64), ("n", ctypes.c_int32), ("max_cells", ctypes.c_int64),
                ("tables_in_hbm", ctypes.c_int32), ("class_count", ctypes.c_int64 * 4), ("class_cells", ctypes.c_int64 * 4),
                ("class_alg_bytes", ctypes.c_int64 * 4), ("all_packed", ctypes.c_int32), ("passes3", ctypes.c_int32), ("items3", ctypes.c_int32)]


class PlanOut(ctypes.Structure):
    _fields_ = [("slices", ctypes.c_int32 * 4), ("ranged", ctypes.c_int32), ("passes", ctypes.c_int32),
                ("cluster", ctypes.c_int32)]


def plan_slices(N: int, n: int, families, sm_count: int = 148, tables_in_hbm: bool = False, all_packed: bool = False) -> dict:
    """The launch plan the library would use for one batch of new families (host arithmetic only,
    works without a GPU).  ``families`` = iterable of (k, cells): number of parents and q*r."""
    bounds = (2048, 12288, 49152)
    pin = PlanIn(sm_count=sm_count, N=N, n=n, max_cells=0, tables_in_hbm=int(tables_in_hbm), all_packed=int(all_packed))
    for k, cells in families:
        cls = sum(cells > b for b in bounds)
        pin.class_count[cls] += 1
        pin.class_cells[cls] += cells
        pin.class_alg_bytes[cls] += (k + 1) * N + 4 * cells
        pin.max_cells = max(pin.max_cells, cells)
    out = PlanOut()
    L = lib()
    L.bic_plan_slices.argtypes = [ctypes.POINTER(PlanIn), ctypes.POINTER(PlanOut)]
    L.bic_plan_slices.restype = ctypes.c_int
    rc = L.bic_plan_slices(ctypes.byref(pin), ctypes.byref(out))
    if rc != 0:
        raise BicError(rc, "bic_plan_slices: bad argument")
    return {"slices": list(out.slices), "ranged": bool(out.ranged), "passes": int(out.passes), "cluster": int(out.cluster)}


def range_plan(cells: int, k: int, rad0: int, counters16: bool = False) -> dict:
    """Sub-ranges of one class-3 family (host arithmetic only, works without a GPU): cells per
    sub-range, passes over the rows, states of the first parent per pass (0: cut by cell index)."""
    span, passes, ns = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
    L = lib()
    L.bic_range_plan.argtypes = [ctypes.c_uint32, ctypes.c_int32, ctypes.c_uint32, ctypes.c_int32,
                                 ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    L.bic_range_plan.restype = ctypes.c_int
    rc = L.bic_range_plan(cells, k, rad0, int(counters16), ctypes.byref(span), ctypes.byref(passes), ctypes.byref(ns))
    if rc != 0:
        raise BicError(rc, "bic_range_plan: bad argument")
    return {"span": span.value, "passes": passes.value, "states_per_pass": ns.value}


class Profile(ctypes.Structure):
    _fields_ = [("count_ms", ctypes.c_double), ("count_launches", ctypes.c_int64),
                ("kernel_launches", ctypes.c_int64), ("families_counted", ctypes.c_int64),
                ("rows_counted", ctypes.c_int64), ("alg_bytes", ctypes.c_int64),
                ("class_ms", ctypes.c_double * 4), ("class_launches", ctypes.c_int64 * 4),
                ("class_families", ctypes.c_int64 * 4), ("class_alg_bytes", ctypes.c_int64 * 4),
                ("families_derived", ctypes.c_int64), ("exchange_ms", ctypes.c_double),
                ("exchange_bytes", ctypes.c_int64), ("exchange_fused", ctypes.c_int64), ("exchange_nccl", ctypes.c_int64)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/libbicgpu.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "bicgpu.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        cmd = ["make", "-C", CSRC] + (["-B"] if force else []) + ["libbicgpu.so"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            print(res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError("building libbicgpu.so failed")
    return LIB_PATH


_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
            "dags_vae_search_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.bic_version.restype = ctypes.c_int
    L.bic_build_info.restype = ctypes.c_char_p
    L.bic_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    L.bic_destroy.argtypes = [vp]
    L.bic_last_error.argtypes = [vp]
    L.bic_last_error.restype = ctypes.c_char_p
    L.bic_set_stream.argtypes = [vp, vp]
    L.bic_wait_stream.argtypes = [vp, vp]
    L.bic_dataset_fingerprint.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    L.bic_score_dags_wire16.argtypes = [vp, vp, vp, i32, i64, ctypes.c_int, vp, ctypes.POINTER(i64), ctypes.c_int]
    L.bic_sync.argtypes = [vp]
    L.bic_set_iss.argtypes = [vp, ctypes.c_double]
    L.bic_set_dataset.argtypes = [vp, vp, i64, i32, i64, vp, ctypes.c_int]
    L.bic_count_families.argtypes = [vp, vp, vp, vp, i64, vp, vp, ctypes.c_int]
    L.bic_score_families.argtypes = [vp, vp, vp, vp, i64, ctypes.c_int, vp, ctypes.c_int]
    L.bic_score_dags_adj.argtypes = [vp, vp, i64, ctypes.c_int, vp, ctypes.POINTER(i64), ctypes.c_int]
    L.bic_score_dags_csr.argtypes = [vp, vp, vp, i64, ctypes.c_int, vp, ctypes.POINTER(i64), ctypes.c_int]
    L.bic_score_dags_wire.argtypes = [vp, vp, vp, i64, ctypes.c_int, vp, ctypes.POINTER(i64), ctypes.c_int]
    L.bic_cache_clear.argtypes = [vp]
    L.bic_cache_reserve.argtypes = [vp, i64]
    L.bic_cache_stats.argtypes = [vp, ctypes.POINTER(CacheStats)]
    L.bic_cache_export.argtypes = [vp, vp, vp, vp, i64, ctypes.POINTER(i64), ctypes.POINTER(i32)]
    L.bic_cache_import.argtypes = [vp, vp, vp, vp, i64, i32]
    L.bic_profile_enable.argtypes = [vp, ctypes.c_int]
    L.bic_profile_reset.argtypes = [vp]
    L.bic_profile_get.argtypes = [vp, ctypes.POINTER(Profile)]
    L.bic_comm_unique_id.argtypes = [vp]
    L.bic_comm_init.argtypes = [vp, vp, ctypes.c_int, ctypes.c_int]
    L.bic_comm_destroy.argtypes = [vp]
    L.bic_comm_mode.argtypes = [vp, ctypes.c_int]
    for name in SYMBOLS:
        if name not in ("bic_last_error", "bic_build_info"):
            getattr(L, name).restype = ctypes.c_int
    _LIB = L
    return L


def build_info() -> str:
    """Provenance of the loaded library (ABI version, arch, compiler, build time) and its path."""
    return lib().bic_build_info().decode() + ", " + os.path.relpath(LIB_PATH, os.path.dirname(_HERE))


def check(ctx, status: int) -> None:
    if status != BIC_OK:
        msg = lib().bic_last_error(ctx)
        raise BicError(status, msg.decode() if msg else "")
