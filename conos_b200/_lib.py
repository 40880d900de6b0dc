"""ctypes binding of libconosgpu.so (include/conos_gpu.h).

There is no CPU fallback: if the shared library is missing, or no B200 is visible when a context is
requested, this module raises.  Build the library with ``python __graft_entry__.py`` (or
``make -C conos_b200/csrc``).
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import sys
import weakref
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libconosgpu.so")

METRICS = {"angular": 0, "L2": 1}
MATCHING = {"mNN": 0, "NN": 1}
SPACES = {"PCA": 0, "CPCA": 1, "CCA": 2}

c_int_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)
c_dbl_p = C.POINTER(C.c_double)
c_flt_p = C.POINTER(C.c_float)


class CgSample(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_genes", C.c_int32), ("p", c_int_p), ("i", c_int_p), ("x", c_dbl_p),
                ("var_v", c_dbl_p), ("var_qv", c_dbl_p), ("pca", c_dbl_p), ("n_pcs", C.c_int32),
                ("pca_ld", C.c_int32)]


class CgPair(C.Structure):
    _fields_ = [("a", C.c_int32), ("b", C.c_int32), ("n_genes", C.c_int32), ("genes_a", c_int_p),
                ("genes_b", c_int_p), ("rot", c_dbl_p), ("rot_rows", C.c_int32), ("rot_cols", C.c_int32),
                ("u", c_dbl_p), ("v", c_dbl_p), ("rows_u", c_int_p), ("rows_v", c_int_p), ("n_u", C.c_int32),
                ("n_v", C.c_int32), ("uv_cols", C.c_int32), ("k", C.c_int32)]


class CgParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("k1", C.c_int32), ("k_self", C.c_int32), ("ncomps", C.c_int32),
                ("space", C.c_int32), ("matching", C.c_int32), ("metric", C.c_int32), ("var_scale", C.c_int32),
                ("common_centering", C.c_int32), ("score_component_variance", C.c_int32),
                ("k_self_weight", C.c_double), ("l2_sigma", C.c_double), ("cor_base", C.c_double),
                ("min_similarity", C.c_double), ("cpca_maxit", C.c_int32), ("cpca_tol", C.c_double)]


# name -> (restype, argtypes); every symbol include/conos_gpu.h declares
_VP = C.c_void_p
SIGNATURES = {
    "cg_init": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "cg_shutdown": (None, [_VP]),
    "cg_last_error": (C.c_char_p, [_VP]),
    "cg_device_count": (C.c_int, []),
    "cg_kernel_launches": (C.c_uint64, [_VP]),
    "cg_knn_timing": (C.c_int, [_VP, C.c_int]),
    "cg_knn_timing_read": (C.c_int, [_VP, c_dbl_p, C.POINTER(C.c_uint64), c_dbl_p, c_dbl_p]),
    "cg_cross_knn": (C.c_int, [_VP, c_dbl_p, C.c_int, C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                               c_int_p, c_flt_p, c_int_p, c_flt_p]),
    "cg_self_knn": (C.c_int, [_VP, c_dbl_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_int_p, c_flt_p]),
    "cg_neighbor_matrix": (C.c_int, [_VP, c_dbl_p, C.c_int, C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, c_i64_p]),
    "cg_fetch_coo": (C.c_int, [_VP, c_int_p, c_int_p, c_dbl_p]),
    "cg_panel_create": (C.c_int, [_VP, C.POINTER(CgSample), C.c_int, C.POINTER(_VP)]),
    "cg_panel_free": (None, [_VP]),
    "cg_panel_offsets": (C.c_int, [_VP, c_int_p]),
    "cg_panel_drop_cache": (C.c_int, [_VP]),
    "cg_params_default": (None, [C.POINTER(CgParams)]),
    "cg_pairs_edges": (C.c_int, [_VP, _VP, C.POINTER(CgPair), C.c_int, C.POINTER(CgParams), C.POINTER(_VP), c_i64_p]),
    "cg_pair_reduction": (C.c_int, [_VP, C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p]),
    "cg_pair_cca": (C.c_int, [_VP, C.c_int, c_int_p, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_int_p, c_int_p, c_dbl_p]),
    "cg_pair_reductions_all": (C.c_int, [_VP, c_int_p, c_int_p, c_i64_p, c_dbl_p, C.c_int64]),
    "cg_pair_cpca": (C.c_int, [_VP, C.c_int, c_int_p, c_dbl_p, c_dbl_p]),
    "cg_pair_cca_loadings": (C.c_int, [_VP, C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "cg_keep_projections": (C.c_int, [_VP, C.c_int]),
    "cg_pair_projection": (C.c_int, [_VP, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p]),
    "cg_local_edges": (C.c_int, [_VP, _VP, c_int_p, C.c_int, C.POINTER(CgParams), C.POINTER(_VP), c_i64_p]),
    "cg_edges_count": (C.c_int64, [_VP]),
    "cg_edges_count_type": (C.c_int64, [_VP, _VP, C.c_int32]),
    "cg_edges_fetch": (C.c_int, [_VP, _VP, c_int_p, c_int_p, c_dbl_p, c_int_p]),
    "cg_edges_append_host": (C.c_int, [_VP, C.POINTER(_VP), c_int_p, c_int_p, c_dbl_p, c_int_p, C.c_int64]),
    "cg_edges_append_device": (C.c_int, [_VP, C.POINTER(_VP), _VP, _VP, _VP, _VP, C.c_int64]),
    "cg_set_option": (C.c_int, [_VP, C.c_char_p, C.c_double]),
    "cg_get_stat": (C.c_int, [_VP, C.c_char_p, C.POINTER(C.c_double)]),
    "cg_profile": (C.c_int, [_VP, C.c_int]),
    "cg_profile_dump": (C.c_int, [_VP, C.c_char_p, C.c_int]),
    "cg_timer_start": (C.c_int, [_VP]),
    "cg_timer_stop": (C.c_int, [_VP, c_dbl_p]),
    "cg_edges_device_ptrs": (C.c_int, [_VP, C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP), C.POINTER(_VP)]),
    "cg_edges_reserve": (C.c_int, [_VP, C.POINTER(_VP), C.c_int64]),
    "cg_edges_set_count": (C.c_int, [_VP, C.c_int64]),
    "cg_edges_free": (None, [_VP]),
    "cg_comm_unique_id": (C.c_int, [_VP]),
    "cg_comm_init": (C.c_int, [_VP, C.c_int, C.c_int, _VP]),
    "cg_comm_rank": (C.c_int, [_VP, c_int_p, c_int_p]),
    "cg_partition_pairs": (C.c_int, [c_int_p, c_int_p, c_dbl_p, C.c_int, c_dbl_p, C.c_int, C.c_int, c_int_p]),
    "cg_job_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "cg_job_free": (None, [_VP]),
    "cg_job_devices": (C.c_int, [_VP]),
    "cg_job_context": (_VP, [_VP, C.c_int]),
    "cg_job_last_error": (C.c_char_p, [_VP]),
    "cg_job_build_graph": (C.c_int, [_VP, C.POINTER(CgSample), C.c_int, C.POINTER(CgPair), C.c_int,
                                     C.POINTER(CgParams), C.POINTER(_VP)]),
    "cg_job_pair_location": (C.c_int, [_VP, C.c_int, c_int_p, c_int_p]),
    "cg_panel_share_grams": (C.c_int, [_VP, _VP, c_int_p, c_int_p]),
    "cg_partition_units": (C.c_int, [c_dbl_p, C.c_int, C.c_int, c_dbl_p, c_int_p]),
    "cg_edges_allgatherv": (C.c_int, [_VP, C.POINTER(_VP), c_int_p, c_i64_p, C.c_int]),
    "cg_assemble_graph": (C.c_int, [_VP, _VP, C.c_int32, C.POINTER(_VP)]),
    "cg_graph_sizes": (C.c_int, [_VP, c_int_p, c_i64_p]),
    "cg_graph_fetch": (C.c_int, [_VP, _VP, c_int_p, c_int_p, c_int_p, c_dbl_p, c_int_p]),
    "cg_graph_fetch_adjacency": (C.c_int, [_VP, _VP, c_i64_p, c_int_p, c_dbl_p]),
    "cg_graph_balance": (C.c_int, [_VP, _VP, c_int_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    "cg_graph_free": (None, [_VP]),
    "cg_spcov": (C.c_int, [_VP, C.c_int, C.c_int, c_int_p, c_int_p, c_dbl_p, c_dbl_p, c_dbl_p]),
    "cg_cpca": (C.c_int, [_VP, c_dbl_p, C.c_int, C.c_int, c_dbl_p, C.c_int, C.c_int, C.c_double, c_dbl_p, c_dbl_p,
                          c_dbl_p, c_dbl_p]),
    "cg_pare_down_hub_edges": (C.c_int, [_VP, C.c_int, c_int_p, c_int_p, c_dbl_p, c_int_p, C.c_int, C.c_int]),
    "cg_sum_weight_matrix": (C.c_int, [_VP, c_dbl_p, c_int_p, c_int_p, C.c_int64, c_int_p, C.c_int, C.c_int, C.c_int,
                                       c_dbl_p]),
    "cg_adjust_weights": (C.c_int, [_VP, c_dbl_p, c_int_p, c_int_p, C.c_int64, c_int_p, C.c_int, C.c_int, c_dbl_p]),
    "cg_reference_wij": (C.c_int, [_VP, c_int_p, c_int_p, c_dbl_p, C.c_int64, C.c_double, c_int_p, c_int_p, c_dbl_p,
                                   c_i64_p]),
    "cg_smooth_matrix": (C.c_int, [_VP, _VP, c_int_p, c_int_p, c_dbl_p, C.c_int64, C.c_int32, c_dbl_p, C.c_int32,
                                   C.POINTER(C.c_uint8), C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32,
                                   c_int_p, c_dbl_p]),
    "cg_propagate_labels": (C.c_int, [_VP, _VP, c_int_p, c_int_p, c_dbl_p, C.c_int64, C.c_int32, c_int_p, C.c_int32,
                                      C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32, c_dbl_p, c_int_p,
                                      c_dbl_p]),
}

_lib: Optional[C.CDLL] = None


class ConosGpuError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libconosgpu.so and bind every declared symbol (raises if the library or a symbol is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ConosGpuError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built (run `python __graft_entry__.py`); "
            "conos_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ctype))


class Context:
    """One cg_context per process and device."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = _VP()
        rc = self.lib.cg_init(device, C.byref(h))
        if rc != 0:
            msg = self.lib.cg_last_error(None)
            raise ConosGpuError(msg.decode() if msg else "cg_init failed")
        self.h = h
        self.device = device
        self._comm = None
        _all_ctx.add(self)

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.cg_last_error(self.h)
            raise ConosGpuError(msg.decode() if msg else f"libconosgpu error {rc}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.cg_shutdown(self.h)
            self.h = None

    def kernel_launches(self) -> int:
        return int(self.lib.cg_kernel_launches(self.h))

    def stat(self, name: str) -> float:
        """Counter of the library context ("knn_redo_blocks", "eig_redo_pairs", "kernel_launches")."""
        v = C.c_double(0.0)
        if self.lib.cg_get_stat(self.h, name.encode(), C.byref(v)) != 0:
            raise KeyError(name)
        return v.value

    # ---- multi-GPU (one context per GPU; the communicator lives inside the library)
    def comm_init(self, rank: int, world: int, group=None):
        """NCCL communicator of this context.  The 128-byte id is created on rank 0 and travels through
        torch.distributed (any backend) -- plumbing only: every collective of the data path runs inside the library."""
        if getattr(self, "_comm", None) == (rank, world):
            return
        ident = (C.c_ubyte * 128)()
        if world > 1:
            import torch
            import torch.distributed as dist
            if rank == 0:
                self.check(self.lib.cg_comm_unique_id(ident))
            dev = torch.device("cuda", self.device) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.tensor(list(bytes(ident)), dtype=torch.uint8, device=dev)
            dist.broadcast(t, src=0, group=group)
            ident = (C.c_ubyte * 128)(*t.cpu().tolist())
        self.check(self.lib.cg_comm_init(self.h, rank, world, ident))
        self._comm = (rank, world)

    def __del__(self):
        # never during interpreter teardown (see _shutdown_all)
        try:
            if not sys.is_finalizing():
                self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class BorrowedContext:
    """A cg_context owned by somebody else (a device of a cg_job): same calling convention, never shut down here."""

    def __init__(self, lib, h, device):
        self.lib, self.h, self.device = lib, _VP(h), device

    def check(self, rc: int):
        if rc != 0:
            msg = self.lib.cg_last_error(self.h)
            raise ConosGpuError(msg.decode() if msg else f"libconosgpu error {rc}")


class Job:
    """cg_job: one buildGraph over n_devices GPUs from this process (one host thread per device inside the library)."""

    def __init__(self, n_devices: int):
        self.lib = load()
        h = _VP()
        if self.lib.cg_job_create(n_devices, C.byref(h)) != 0:
            msg = self.lib.cg_last_error(None)
            raise ConosGpuError(msg.decode() if msg else "cg_job_create failed")
        self.h = h
        self.n_devices = n_devices
        self.contexts = [BorrowedContext(self.lib, self.lib.cg_job_context(h, d), d) for d in range(n_devices)]
        _all_jobs.add(self)

    def close(self):
        if getattr(self, "h", None):
            self.lib.cg_job_free(self.h)
            self.h = None

    def __del__(self):
        try:
            if not sys.is_finalizing():
                self.close()
        except Exception:
            pass


_default_ctx = {}
_all_ctx = weakref.WeakSet()
_all_jobs = weakref.WeakSet()


def default_context(device: int = 0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


@atexit.register
def _shutdown_all():
    """Contexts are shut down by an atexit hook, i.e. while the interpreter, the CUDA runtime and torch are all still
    intact -- never from a destructor during interpreter teardown (the order in which modules and extension state go
    away there is arbitrary).  R's equivalent is the package's .onUnload -> cg_shutdown."""
    for j in list(_all_jobs):
        try:
            j.close()
        except Exception:
            pass
    for c in list(_all_ctx):
        try:
            c.close()
        except Exception:
            pass
    _default_ctx.clear()
