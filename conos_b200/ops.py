"""Stage-level operators with the reference's names and argument meaning, running on the B200.

Each function is a thin marshalling layer over one C-ABI entry point of libconosgpu.so
(include/conos_gpu.h); no arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import METRICS, MATCHING, ptr


def _ctx(ctx=None):
    return ctx if ctx is not None else _lib.default_context()


def _metric(metric: str) -> int:
    if metric not in METRICS:
        raise ValueError("only the following distance metrics are currently supported: ['L2' 'angular']")
    return METRICS[metric]


def _colmajor(x) -> np.ndarray:
    return np.asfortranarray(np.asarray(x, dtype=np.float64))


def crossKnn(mA, mB, k: int, metric: str = "angular", ctx=None, both: bool = False):
    """N2R::crossKnn(mA, mB, k, indexType=metric) with exact search (R/conos.R:851-855): for every row of
    mA its k nearest rows of mB.  Returns (idx [nA,k] int32, dist [nA,k] float32); with both=True also the
    reverse direction from the same pass: (idx_ab, dist_ab, idx_ba, dist_ba)."""
    c = _ctx(ctx)
    A, B = _colmajor(mA), _colmajor(mB)
    if A.shape[1] != B.shape[1]:
        raise ValueError("mA and mB must have the same number of columns")
    nA, d = A.shape
    nB = B.shape[0]
    iab = np.empty((nA, k), np.int32)
    dab = np.empty((nA, k), np.float32)
    iba = np.empty((nB, k), np.int32) if both else None
    dba = np.empty((nB, k), np.float32) if both else None
    c.check(c.lib.cg_cross_knn(c.h, ptr(A, C.c_double), nA, max(nA, 1), ptr(B, C.c_double), nB, max(nB, 1), d, k,
                               _metric(metric), ptr(iab, C.c_int32), ptr(dab, C.c_float), ptr(iba, C.c_int32),
                               ptr(dba, C.c_float)))
    return (iab, dab, iba, dba) if both else (iab, dab)


def Knn(m, k: int, metric: str = "angular", ctx=None):
    """N2R::Knn(m, k, indexType=metric) with exact search (R/conos.R:595): the point itself is included."""
    c = _ctx(ctx)
    Z = _colmajor(m)
    n, d = Z.shape
    idx = np.empty((n, k), np.int32)
    dist = np.empty((n, k), np.float32)
    c.check(c.lib.cg_self_knn(c.h, ptr(Z, C.c_double), n, max(n, 1), d, k, _metric(metric), ptr(idx, C.c_int32),
                              ptr(dist, C.c_float)))
    return idx, dist


def getNeighborMatrix(p1, p2, k: int, k1: Optional[int] = None, matching: str = "mNN", metric: str = "angular",
                      l2_sigma: float = 1e5, cor_base: float = 1.0, min_similarity: float = 1e-5, ctx=None):
    """getNeighborMatrix (R/conos.R:848-886): triplets (i, j, x) of the n1 x n2 similarity matrix in CSC order."""
    c = _ctx(ctx)
    if matching not in MATCHING:
        raise ValueError("Unrecognized type of NN matching:" + str(matching))
    if k1 is None:
        k1 = k
    A, B = _colmajor(p1), _colmajor(p2)
    nnz = C.c_int64(0)
    c.check(c.lib.cg_neighbor_matrix(c.h, ptr(A, C.c_double), A.shape[0], max(A.shape[0], 1), ptr(B, C.c_double),
                                     B.shape[0], max(B.shape[0], 1), A.shape[1], k, k1, MATCHING[matching],
                                     _metric(metric), l2_sigma, cor_base, min_similarity, C.byref(nnz)))
    n = nnz.value
    i = np.empty(n, np.int32)
    j = np.empty(n, np.int32)
    x = np.empty(n, np.float64)
    c.check(c.lib.cg_fetch_coo(c.h, ptr(i, C.c_int32), ptr(j, C.c_int32), ptr(x, C.c_double)))
    return i, j, x


def spcov(m: sp.spmatrix, cm: np.ndarray, ctx=None) -> np.ndarray:
    """spcov(m, cm) (src/spcov.cpp:14-19): dense G x G covariance of a sparse cells x genes matrix."""
    c = _ctx(ctx)
    m = sp.csc_matrix(m)
    m.sort_indices()
    n, G = m.shape
    p = np.ascontiguousarray(m.indptr, np.int32)
    i = np.ascontiguousarray(m.indices, np.int32)
    x = np.ascontiguousarray(m.data, np.float64)
    cm = np.ascontiguousarray(cm, np.float64)
    out = np.empty((G, G), np.float64, order="F")
    c.check(c.lib.cg_spcov(c.h, n, G, ptr(p, C.c_int32), ptr(i, C.c_int32), ptr(x, C.c_double), ptr(cm, C.c_double),
                           ptr(out, C.c_double)))
    return out


def cpcaF(cov: np.ndarray, ng, ncomp: int = 10, maxit: int = 1000, tol: float = 1e-6, eigenvR=None, ctx=None):
    """cpcaF(cov p x p x k, ng, ncomp, maxit, tol, eigenvR) (src/cpca.cpp:16-101) -> dict(D, CPC, niter)."""
    c = _ctx(ctx)
    if eigenvR is None:
        raise ValueError("the device cpcaF needs eigenvR (conos always supplies it, R/conos.R:185-187)")
    cov = np.asfortranarray(cov, np.float64)
    p, _, k = cov.shape
    ng = np.ascontiguousarray(ng, np.float64)
    ev = np.asfortranarray(np.asarray(eigenvR, np.float64)[:, :ncomp])
    D = np.zeros((ncomp, k), order="F")
    CPC = np.zeros((p, ncomp), order="F")
    niter = np.zeros(ncomp)
    c.check(c.lib.cg_cpca(c.h, ptr(cov, C.c_double), p, k, ptr(ng, C.c_double), ncomp, maxit, tol,
                          ptr(ev, C.c_double), ptr(D, C.c_double), ptr(CPC, C.c_double), ptr(niter, C.c_double)))
    return {"D": D, "CPC": CPC, "niter": niter}


def pareDownHubEdges(sY: sp.csc_matrix, rowN: np.ndarray, k: int, klow: int = -1, ctx=None) -> np.ndarray:
    """pareDownHubEdges(sY, rowN, k, klow) (src/edgeFilter.cpp:23-60): returns the new @x; rowN is mutated."""
    c = _ctx(ctx)
    sY = sp.csc_matrix(sY)
    p = np.ascontiguousarray(sY.indptr, np.int32)
    i = np.ascontiguousarray(sY.indices, np.int32)
    x = np.ascontiguousarray(sY.data, np.float64).copy()
    if rowN.dtype != np.int32 or not rowN.flags.c_contiguous:
        raise ValueError("rowN must be a contiguous int32 array (it is updated in place)")
    c.check(c.lib.cg_pare_down_hub_edges(c.h, sY.shape[1], ptr(p, C.c_int32), ptr(i, C.c_int32), ptr(x, C.c_double),
                                         ptr(rowN, C.c_int32), k, klow))
    return x


def getSumWeightMatrix(weights, row_inds, col_inds, factor_levels, normalize: bool = True, ctx=None) -> np.ndarray:
    """getSumWeightMatrix (src/edge_rebalancing.cpp:14-36): n_cells x n_levels."""
    c = _ctx(ctx)
    w = np.ascontiguousarray(weights, np.float64)
    ri = np.ascontiguousarray(row_inds, np.int32)
    ci = np.ascontiguousarray(col_inds, np.int32)
    fl = np.ascontiguousarray(factor_levels, np.int32)
    n_cells, n_levels = len(fl), int(fl.max())
    m = np.zeros((n_cells, n_levels), order="F")
    c.check(c.lib.cg_sum_weight_matrix(c.h, ptr(w, C.c_double), ptr(ri, C.c_int32), ptr(ci, C.c_int32), len(w),
                                       ptr(fl, C.c_int32), n_cells, n_levels, int(normalize), ptr(m, C.c_double)))
    return m


def adjustWeightsByCellBalancingC(weights, row_inds, col_inds, factor_levels, dividers, ctx=None) -> np.ndarray:
    """adjustWeightsByCellBalancingC (src/edge_rebalancing.cpp:39-48)."""
    c = _ctx(ctx)
    w = np.ascontiguousarray(weights, np.float64).copy()
    ri = np.ascontiguousarray(row_inds, np.int32)
    ci = np.ascontiguousarray(col_inds, np.int32)
    fl = np.ascontiguousarray(factor_levels, np.int32)
    dv = np.asfortranarray(dividers, np.float64)
    c.check(c.lib.cg_adjust_weights(c.h, ptr(w, C.c_double), ptr(ri, C.c_int32), ptr(ci, C.c_int32), len(w),
                                    ptr(fl, C.c_int32), len(fl), dv.shape[1], ptr(dv, C.c_double)))
    return w


def referenceWij(i, j, d, perplexity: float, n: Optional[int] = None, ctx=None) -> sp.csc_matrix:
    """referenceWij(i, j, d, threads, perplexity) (src/edgeweights.cpp:160-184)."""
    c = _ctx(ctx)
    i = np.ascontiguousarray(i, np.int32)
    j = np.ascontiguousarray(j, np.int32)
    d = np.ascontiguousarray(d, np.float64)
    of = np.empty(2 * len(i) + 1, np.int32)
    ot = np.empty(2 * len(i) + 1, np.int32)
    ow = np.empty(2 * len(i) + 1, np.float64)
    no = C.c_int64(0)
    c.check(c.lib.cg_reference_wij(c.h, ptr(i, C.c_int32), ptr(j, C.c_int32), ptr(d, C.c_double), len(i), perplexity,
                                   ptr(of, C.c_int32), ptr(ot, C.c_int32), ptr(ow, C.c_double), C.byref(no)))
    ne = no.value
    if n is None:
        n = int(i[-1]) + 1 if len(i) else 0
    return sp.coo_matrix((ow[:ne], (of[:ne], ot[:ne])), shape=(n, n)).tocsc()


def buildWijMatrix(x: sp.spmatrix, perplexity: float = 50, ctx=None) -> sp.csc_matrix:
    """buildWijMatrix.CsparseMatrix (R/largeVis.R:30-34)."""
    a = sp.csc_matrix(x)
    a.sort_indices()
    is_ = np.repeat(np.arange(a.shape[1], dtype=np.int32), np.diff(a.indptr))
    return referenceWij(is_, a.indices, a.data ** 2, perplexity, a.shape[0], ctx=ctx)


def smooth_count_matrix(edge_from, edge_to, edge_weights, count_matrix, is_label_fixed=None, max_n_iters: int = 10,
                        diffusion_fading: float = 1.0, diffusion_fading_const: float = 0.1, tol: float = 1e-3,
                        normalize: bool = False, ctx=None, graph=None, return_info: bool = False):
    """smooth_count_matrix (src/propagate_labels.cpp:177-222): vertices are 0-based row indices of `count_matrix`
    (n_vertices x n_cols).  `graph`: a device graph handle instead of the host edge list."""
    c = _ctx(ctx)
    m = np.array(count_matrix, dtype=np.float64, order="F", copy=True)
    if m.ndim != 2:
        raise ValueError("count_matrix must be a matrix")
    ef = np.ascontiguousarray(edge_from, np.int32) if edge_from is not None else np.zeros(0, np.int32)
    et = np.ascontiguousarray(edge_to, np.int32) if edge_to is not None else np.zeros(0, np.int32)
    w = np.ascontiguousarray(edge_weights, np.float64) if edge_weights is not None else np.zeros(0)
    if len(ef) != len(w) or len(et) != len(w):
        raise ValueError("Size of edge_verts must match size of edge_weights")
    fx = None
    if is_label_fixed is not None and len(is_label_fixed):
        fx = np.ascontiguousarray(is_label_fixed, np.uint8)
        if len(fx) != m.shape[0]:
            raise ValueError("is_label_fixed must have one entry per vertex")
    it, nrm = C.c_int32(0), C.c_double(0)
    c.check(c.lib.cg_smooth_matrix(c.h, graph, ptr(ef, C.c_int32), ptr(et, C.c_int32), ptr(w, C.c_double), len(w),
                                   m.shape[0], ptr(m, C.c_double), m.shape[1],
                                   None if fx is None else ptr(fx, C.c_uint8), int(max_n_iters), float(diffusion_fading),
                                   float(diffusion_fading_const), float(tol), int(bool(normalize)), C.byref(it),
                                   C.byref(nrm)))
    return (m, it.value, nrm.value) if return_info else m


def propagate_labels(edge_from, edge_to, edge_weights, n_vertices: int, label_of_vertex, n_labels: int,
                     max_n_iters: int = 10, diffusion_fading: float = 10.0, diffusion_fading_const: float = 0.5,
                     tol: float = 5e-3, fixed_initial_labels: bool = False, ctx=None, graph=None,
                     return_info: bool = False):
    """propagate_labels (src/propagate_labels.cpp:67-118) on integer ids: label_of_vertex[v] in [0, n_labels) or -1."""
    c = _ctx(ctx)
    ef = np.ascontiguousarray(edge_from, np.int32) if edge_from is not None else np.zeros(0, np.int32)
    et = np.ascontiguousarray(edge_to, np.int32) if edge_to is not None else np.zeros(0, np.int32)
    w = np.ascontiguousarray(edge_weights, np.float64) if edge_weights is not None else np.zeros(0)
    if len(ef) != len(w) or len(et) != len(w):
        raise ValueError("Incorrect dimension of input vectors")
    lab = np.ascontiguousarray(label_of_vertex, np.int32)
    if len(lab) != n_vertices:
        raise ValueError("one label id per vertex expected")
    out = np.zeros((n_vertices, n_labels), dtype=np.float64, order="F")
    it, nrm = C.c_int32(0), C.c_double(0)
    c.check(c.lib.cg_propagate_labels(c.h, graph, ptr(ef, C.c_int32), ptr(et, C.c_int32), ptr(w, C.c_double), len(w),
                                      int(n_vertices), ptr(lab, C.c_int32), int(n_labels), int(max_n_iters),
                                      float(diffusion_fading), float(diffusion_fading_const), float(tol),
                                      int(bool(fixed_initial_labels)), ptr(out, C.c_double), C.byref(it), C.byref(nrm)))
    return (out, it.value, nrm.value) if return_info else out
