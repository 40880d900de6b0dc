"""Pair sharding and the edge-table allgatherv for one-process-per-GPU runs (torch.distributed plumbing).

The sample-pair list is cut into contiguous ranges balanced by cells_a x cells_b, so the gathered table is
simply the concatenation of the ranks' tables in rank order == the reference's combn order
(R/conclass.R:1058, 224-317).  The only exchange step of the path is this allgatherv (NCCL on the GPU,
gloo in the CPU tests); there is no other data-path collective.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def partition_contiguous(costs: Sequence[float], n_parts: int) -> List[Tuple[int, int]]:
    """Cut range(len(costs)) into n_parts contiguous ranges with near-equal cost sums.
    Boundaries are placed where the running sum crosses i/n of the total (deterministic, O(n))."""
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if n_parts <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (max(n_parts, 1) - 1)
    cum = np.concatenate([[0.0], np.cumsum(costs)])
    total = cum[-1]
    bounds = [0]
    for i in range(1, n_parts):
        target = total * i / n_parts
        # first index whose prefix sum reaches the target, but keep ranges non-decreasing
        b = int(np.searchsorted(cum, target, side="left"))
        # choose the closer of b-1 / b
        if b > 0 and abs(cum[b - 1] - target) <= abs(cum[min(b, n)] - target):
            b -= 1
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(n_parts)]


def allgather_columns(cols, group=None):
    """allgatherv of equally long 1-D tensors (one per column of the edge table): every rank receives the
    concatenation over ranks, in rank order.  Works for CUDA tensors (NCCL) and CPU tensors (gloo).
    Returns (gathered columns, counts per rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = int(cols[0].numel())
    dev = cols[0].device
    cnt = torch.tensor([n], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, cnt, group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    out = []
    for c in cols:
        pad = torch.zeros(mx, dtype=c.dtype, device=dev)
        pad[:n] = c
        parts = [torch.empty(mx, dtype=c.dtype, device=dev) for _ in range(world)]
        if mx > 0:
            dist.all_gather(parts, pad, group=group)
        out.append(torch.cat([p[:k] for p, k in zip(parts, counts)]) if mx > 0
                   else torch.empty(0, dtype=c.dtype, device=dev))
    return out, counts


class _DevArray:
    """Zero-copy view of library-owned device memory for torch (CUDA array interface v2)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_columns(ctx, edges):
    """torch CUDA tensors aliasing the four columns of a cg_edges table (from, to, w, type)."""
    import ctypes as C
    import torch
    n = int(ctx.lib.cg_edges_count(edges))
    f, t, w, ty = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    ctx.lib.cg_edges_device_ptrs(edges, C.byref(f), C.byref(t), C.byref(w), C.byref(ty))
    dev = torch.device("cuda", ctx.device)
    if n == 0:
        return [torch.empty(0, dtype=torch.int32, device=dev), torch.empty(0, dtype=torch.int32, device=dev),
                torch.empty(0, dtype=torch.float64, device=dev), torch.empty(0, dtype=torch.int32, device=dev)]
    mk = lambda p, ts: torch.as_tensor(_DevArray(p.value, n, ts), device=dev)
    return [mk(f, "<i4"), mk(t, "<i4"), mk(w, "<f8"), mk(ty, "<i4")]


def gather_edges(ctx, edges, group=None):
    """Replace a rank-local cg_edges table by the table gathered over all ranks (rank order). Returns the new
    handle and the per-rank counts."""
    import ctypes as C
    import torch
    cols = device_columns(ctx, edges)
    torch.cuda.synchronize(ctx.device)
    gathered, counts = allgather_columns(cols, group)
    torch.cuda.synchronize(ctx.device)
    new = C.c_void_p()
    n = int(gathered[0].numel())
    ctx.check(ctx.lib.cg_edges_append_device(ctx.h, C.byref(new), C.c_void_p(gathered[0].data_ptr()),
                                             C.c_void_p(gathered[1].data_ptr()), C.c_void_p(gathered[2].data_ptr()),
                                             C.c_void_p(gathered[3].data_ptr()), n))
    ctx.lib.cg_edges_free(edges)
    return new, counts
