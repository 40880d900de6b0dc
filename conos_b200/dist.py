"""Host-side helpers of one-process-per-GPU runs.

Everything on the data path lives in libconosgpu.so: the partition of the pair list (cg_partition_pairs), the NCCL
communicator (cg_comm_init) and the allgatherv of the per-rank edge tables in the reference's row order
(cg_edges_allgatherv; R/conclass.R:224-317 maps over the pairs, :316-321 rbinds the per-pair tables).  This module
holds the thin Python bindings of the two functions that need no GPU, and a NumPy restatement of the gathered
layout that the CPU tests (gloo, world_size 2) check the host logic with.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ptr


def partition_pairs(a: Sequence[int], b: Sequence[int], cost: Sequence[float], world: int,
                    sample_cost: Optional[Sequence[float]] = None, n_samples: Optional[int] = None) -> np.ndarray:
    """owner[q] = rank of pair q (cg_partition_pairs: longest-processing-time greedy with sample locality).
    Deterministic and GPU-free: every rank computes the same table."""
    lib = _lib.load()
    a = np.ascontiguousarray(a, np.int32)
    b = np.ascontiguousarray(b, np.int32)
    cost = np.ascontiguousarray(cost, np.float64)
    if n_samples is None:
        n_samples = int(max(a.max(initial=-1), b.max(initial=-1)) + 1)
    sc = None if sample_cost is None else np.ascontiguousarray(sample_cost, np.float64)
    owner = np.zeros(len(a), np.int32)
    rc = lib.cg_partition_pairs(ptr(a, C.c_int32), ptr(b, C.c_int32), ptr(cost, C.c_double), len(a),
                                ptr(sc, C.c_double), n_samples, world, ptr(owner, C.c_int32))
    if rc != 0:
        raise _lib.ConosGpuError((lib.cg_last_error(None) or b"cg_partition_pairs failed").decode())
    return owner


def gathered_layout(seg_owner: Sequence[int], seg_counts: Sequence[int], world: int):
    """The layout cg_edges_allgatherv builds (comm.cu): every rank's table lands in a staging table rank after rank,
    then segment s moves from staging row src_off[s] to row dst_off[s] of the gathered table (segments ascending).
    Returns (rank_off [world + 1], src_off [n_seg], dst_off [n_seg + 1])."""
    seg_owner = np.asarray(seg_owner, np.int64)
    seg_counts = np.asarray(seg_counts, np.int64)
    rank_total = np.bincount(seg_owner, weights=seg_counts, minlength=world).astype(np.int64)
    rank_off = np.concatenate([[0], np.cumsum(rank_total)])
    dst_off = np.concatenate([[0], np.cumsum(seg_counts)])
    src_off = np.zeros(len(seg_owner), np.int64)
    cursor = rank_off[:-1].copy()
    for s, (o, c) in enumerate(zip(seg_owner, seg_counts)):
        src_off[s] = cursor[o]
        cursor[o] += c
    return rank_off, src_off, dst_off


def permute_gathered(staging: np.ndarray, src_off, dst_off) -> np.ndarray:
    """NumPy restatement of permute_segments_kernel: rows of the staging table in gathered (segment) order."""
    out = np.empty_like(staging)
    for s in range(len(src_off)):
        n = dst_off[s + 1] - dst_off[s]
        out[dst_off[s]:dst_off[s] + n] = staging[src_off[s]:src_off[s] + n]
    return out
