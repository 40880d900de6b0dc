"""ORACLE -- TEST INFRASTRUCTURE ONLY.

NumPy/SciPy restatement of the conos ``buildGraph`` hot path (reference v1.5.3).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import
this module; the product (``conos_b200`` -> ``libconosgpu.so``) never does.

PARITY UNPINNED: the reference's tests hold no golden vectors for this path and the
reference cannot run here (no R; N2R, irlba, Matrix, igraph are not vendored).  Every function cites the
reference lines it restates; behaviour of the un-vendored packages is restated from their published
semantics and marked [3P].

Loop-heavy pieces (exact kNN in canonical float32, pareDownHubEdges, cpcaF, edge re-balancing,
referenceWij) live in ``conos_oracle.c`` and are reached through ctypes.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libconos_oracle.so")
_lib = None

ANGULAR, L2 = 0, 1


def build(force: bool = False) -> str:
    """Compile conos_oracle.c (gcc + OpenMP) into oracle/_build/."""
    src = os.path.join(_HERE, "conos_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def metric_id(metric: str) -> int:
    if metric == "angular":
        return ANGULAR
    if metric == "L2":
        return L2
    raise ValueError("only the following distance metrics are currently supported: ['L2' 'angular']")


# ----------------------------------------------------------------------------------------------------
# kNN  -- N2R::crossKnn / N2R::Knn call sites R/conos.R:595, 851-855  [3P: n2 float32 arithmetic]
# ----------------------------------------------------------------------------------------------------
def prepare_f32(x: np.ndarray, metric: str = "angular") -> np.ndarray:
    """float32 cast (+ L2 normalisation for 'angular') of an n x d double matrix -> row-major float32."""
    x = np.asfortranarray(np.asarray(x, dtype=np.float64))
    n, d = x.shape
    out = np.empty((n, d), dtype=np.float32)
    lib().oracle_prepare_f32(_p(x, ctypes.c_double), n, d, max(n, 1), metric_id(metric), _p(out, ctypes.c_float))
    return out


def knn_f32(q: np.ndarray, r: np.ndarray, k: int, metric: str = "angular", nthreads: int = 0):
    """Exact kNN of prepared float32 rows: (idx [nq,k] int32, dist [nq,k] float32), order (dist, index)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    nq, d = q.shape
    nr = r.shape[0]
    idx = np.empty((nq, k), dtype=np.int32)
    dist = np.empty((nq, k), dtype=np.float32)
    lib().oracle_knn_f32(_p(q, ctypes.c_float), nq, _p(r, ctypes.c_float), nr, d, k, metric_id(metric),
                         _p(idx, ctypes.c_int), _p(dist, ctypes.c_float), nthreads)
    return idx, dist


def knn_f64(q: np.ndarray, r: np.ndarray, k: int, metric: str = "angular"):
    """FP64 brute force on the raw (un-normalised) rows: ground truth for recall@k."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    r = np.ascontiguousarray(r, dtype=np.float64)
    nq, d = q.shape
    idx = np.empty((nq, k), dtype=np.int32)
    dist = np.empty((nq, k), dtype=np.float64)
    lib().oracle_knn_f64(_p(q, ctypes.c_double), nq, _p(r, ctypes.c_double), r.shape[0], d, k, metric_id(metric),
                         _p(idx, ctypes.c_int), _p(dist, ctypes.c_double))
    return idx, dist


def cross_knn(mA: np.ndarray, mB: np.ndarray, k: int, metric: str = "angular"):
    """N2R::crossKnn(mA, mB, k): index mB, query the rows of mA [3P].  Returns (idx, dist) per row of mA;
    as a sparse matrix this is nrow(mB) x nrow(mA) with entry (neighbour, query) = distance."""
    return knn_f32(prepare_f32(mA, metric), prepare_f32(mB, metric), k, metric)


def self_knn(m: np.ndarray, k: int, metric: str = "angular"):
    """N2R::Knn(m, k): search the sample against itself, the point itself included [3P]."""
    p = prepare_f32(m, metric)
    return knn_f32(p, p, k, metric)


# ----------------------------------------------------------------------------------------------------
# scalars of buildGraph -- R/conclass.R:190-198, 223
# ----------------------------------------------------------------------------------------------------
def r_round(x: float) -> int:
    """R's round(): IEC 60559 half-to-even (Python's round has the same rule)."""
    return int(round(x))


def resolve_k1(k: int, k1: Optional[int], alignment_strength: Optional[float], max_ncells: int):
    """returns (k1, alignment_strength) -- R/conclass.R:190-198."""
    if k1 is None:
        k1 = k
    if alignment_strength is not None:
        a = min(max(alignment_strength, 0.0), 1.0)
        k1 = max(r_round(max_ncells * a ** 2), k)
    else:
        a = 0.0
    if k1 < k:
        raise ValueError("k1 must be >= k")
    return k1, a


def cor_base_of(alignment_strength: float) -> float:
    return 1.0 + min(1.0, alignment_strength * 10.0)  # R/conclass.R:223


def convert_distance_to_similarity(d: np.ndarray, metric: str, l2_sigma: float = 1e5, cor_base: float = 1.0):
    """R/conos.R:767-773 (double arithmetic on the float32-valued distances)."""
    d = np.asarray(d, dtype=np.float64)
    if metric == "angular":
        return np.maximum(0.0, cor_base - d)
    return np.exp(-d / l2_sigma)


# ----------------------------------------------------------------------------------------------------
# samples  -- the fields the path reads from a Pagoda2 object (R/access_wrappers.R:9,34,65,85,157)
# ----------------------------------------------------------------------------------------------------
@dataclass
class Sample:
    counts: sp.csc_matrix            # cells x genes, log-scale expression (r$counts)
    genes: List[str]                 # colnames(counts)
    cells: List[str]                 # rownames(counts)
    varinfo_v: np.ndarray            # misc$varinfo$v   aligned with genes
    varinfo_qv: np.ndarray           # misc$varinfo$qv  aligned with genes
    pca: np.ndarray                  # reductions$PCA   cells x nPcs
    od_genes: List[str] = field(default_factory=list)  # getOdGenes() order (most overdispersed first) [3P]


def common_overdispersed_genes(samples: Sequence[Sample], n_odgenes: int) -> List[str]:
    """R/conos.R:107-117.  sort(table(.), decreasing=TRUE) is a stable sort of the name-sorted table, so
    ties keep name order; names are compared in the C locale here."""
    counts: Dict[str, int] = {}
    for s in samples:
        for g in s.od_genes[:n_odgenes]:
            counts[g] = counts.get(g, 0) + 1
    names = sorted(counts)
    order = sorted(range(len(names)), key=lambda t: -counts[names[t]])
    common = set(samples[0].genes)
    for s in samples[1:]:
        common &= set(s.genes)
    od = [names[t] for t in order if names[t] in common]
    return od[:min(len(od), n_odgenes)]


def scaled_matrices(samples: Sequence[Sample], od_genes: Sequence[str], var_scale: bool = True):
    """R/conos.R:18-46 (scaledMatricesP2, data.type='counts')."""
    cols = []
    for s in samples:
        pos = {g: t for t, g in enumerate(s.genes)}
        cols.append(np.array([pos[g] for g in od_genes], dtype=np.int64))
    if var_scale:
        with np.errstate(divide="ignore", invalid="ignore"):
            lq = np.stack([np.log(s.varinfo_qv[c]) for s, c in zip(samples, cols)], axis=1)
            cqv = np.exp(lq.mean(axis=1))
    out = []
    for s, c in zip(samples, cols):
        x = s.counts[:, c].tocsc().astype(np.float64)
        if var_scale:
            with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
                cgsf = np.sqrt(cqv / np.exp(s.varinfo_v[c]))
            cgsf[~np.isfinite(cgsf)] = 0.0
            x = x.copy()
            x.data = x.data * np.repeat(cgsf, np.diff(x.indptr))
        out.append(x)
    return out


def _col_var(x: sp.spmatrix) -> np.ndarray:
    """apply(x, 2, var) -- unbiased per-column variance."""
    n = x.shape[0]
    m = np.asarray(x.mean(axis=0)).ravel()
    m2 = np.asarray(x.multiply(x).mean(axis=0)).ravel()
    return (m2 - m * m) * n / (n - 1)


def top_right_singular_vectors(x: sp.spmatrix, nv: int) -> np.ndarray:
    """Stand-in for irlba::irlba(x, nv, center=colMeans(x)) [3P]: exact eigenvectors of the centred Gram
    matrix (irlba itself is approximate, tol 1e-5, and sign-ambiguous -- compare subspaces, not vectors)."""
    n, g = x.shape
    cm = np.asarray(x.mean(axis=0)).ravel()
    gram = (x.T @ x).toarray() - n * np.outer(cm, cm)
    w, v = np.linalg.eigh(gram)
    return v[:, ::-1][:, :nv]


def quick_plain_pca(samples: Sequence[Sample], ncomps: int = 30, n_odgenes: int = 2000, var_scale: bool = True,
                    score_component_variance: bool = False):
    """R/conos.R:273-315.  Returns dict(CPC = G x 2*round(ncomps/2), genes, nv) or None (<5 genes)."""
    od = common_overdispersed_genes(samples, n_odgenes)
    if len(od) < 5:
        return None
    sm = scaled_matrices(samples, od, var_scale)
    nper = r_round(ncomps / 2)  # the min(nrow(cm)-1, ...) guard at :287 is vacuous (cm is a vector)
    pcs, nvs = [], []
    for x in sm:
        v = top_right_singular_vectors(x, nper)
        pcs.append(v)
        if score_component_variance:
            cm = np.asarray(x.mean(axis=0)).ravel()
            rot = (x @ v) - (cm @ v)[None, :]
            nvs.append(np.var(rot, axis=0, ddof=1) / _col_var(x).sum())
    n1, n2 = pcs[0].shape[1], pcs[1].shape[1]
    pcj = np.concatenate(pcs, axis=1)
    # interleave <- order(c((1:n1)-0.5, 1:n2))   (R/conos.R:300-304)
    keys = np.concatenate([np.arange(1, n1 + 1) - 0.5, np.arange(1, n2 + 1)])
    pcj = pcj[:, np.argsort(keys, kind="stable")]
    res = {"CPC": pcj, "genes": od}
    if score_component_variance:
        res["nv"] = nvs
    return res


def spcov(x: sp.spmatrix, cm: np.ndarray) -> np.ndarray:
    """src/spcov.cpp:14-19."""
    n = x.shape[0]
    v = (x.T @ x).toarray()
    v -= np.outer(cm, cm) * float(n)
    v /= float(n) - 1.0
    return v


def cpca_f(covs: Sequence[np.ndarray], ng: Sequence[float], ncomp: int, maxit: int, tol: float, eigv: np.ndarray):
    """src/cpca.cpp:16-101 with eigenvR supplied."""
    p = covs[0].shape[0]
    k = len(covs)
    cube = np.asfortranarray(np.stack(covs, axis=2), dtype=np.float64)  # p x p x k, slices contiguous
    ngv = np.asarray(ng, dtype=np.float64)
    ev = np.asfortranarray(eigv[:, :ncomp], dtype=np.float64)
    D = np.zeros((ncomp, k), order="F")
    CPC = np.zeros((p, ncomp), order="F")
    niter = np.zeros(ncomp)
    lib().oracle_cpca(_p(cube, ctypes.c_double), p, k, _p(ngv, ctypes.c_double), ncomp, maxit, ctypes.c_double(tol),
                      _p(ev, ctypes.c_double), _p(D, ctypes.c_double), _p(CPC, ctypes.c_double),
                      _p(niter, ctypes.c_double))
    return {"D": D, "CPC": CPC, "niter": niter}


def cpca_fast(covs, ncells, ncomp=10, maxit=1000, tol=1e-6):
    """R/conos.R:175-191 (use.irlba branch; irlba replaced by exact eigh [3P])."""
    p = covs[0].shape[0]
    ncomp = min(p - 1, ncomp)
    tot = float(sum(ncells))
    S = np.zeros((p, p))
    for c, n in zip(covs, ncells):
        S = S + (n / tot) * c
    w, v = np.linalg.eigh(S)
    ev = v[:, ::-1][:, :ncomp]  # irlba(S)$v: leading singular vectors; S is PSD so these are eigenvectors
    return cpca_f(covs, ncells, ncomp, maxit, tol, ev), ev


def quick_cpca(samples: Sequence[Sample], ncomps: int = 100, n_odgenes: int = 2000, var_scale: bool = True,
               score_component_variance: bool = False):
    """R/conos.R:204-259."""
    od = common_overdispersed_genes(samples, n_odgenes)
    if len(od) < 5:
        return None
    ncomps = min(ncomps, len(od) - 1)
    sm = scaled_matrices(samples, od, var_scale)
    covl = [spcov(x, np.asarray(x.mean(axis=0)).ravel()) for x in sm]
    ncells = [x.shape[0] for x in sm]
    res, ev0 = cpca_fast(covl, ncells, ncomp=ncomps, maxit=500, tol=1e-5)
    res["genes"] = od
    res["init"] = ev0
    if score_component_variance:
        nv = []
        for x in sm:
            cm = np.asarray(x.mean(axis=0)).ravel()
            rot = (x @ res["CPC"]) - (cm @ res["CPC"])[None, :]
            nv.append(np.var(rot, axis=0, ddof=1) / _col_var(x).sum())
        res["nv"] = nv
    return res


def quick_cca(samples: Sequence[Sample], ncomps: int = 100, n_odgenes: int = 2000, var_scale: bool = True):
    """R/conos.R:327-369 (PMA=FALSE).  irlba replaced by a dense SVD [3P]."""
    od = common_overdispersed_genes(samples, n_odgenes)
    if len(od) < 5:
        return None
    ncomps = min(ncomps, len(od) - 1)
    sm = scaled_matrices(samples, od, var_scale)
    keep = [np.asarray(m.sum(axis=1)).ravel() > 0 for m in sm]
    dm = [m[kp].toarray() for m, kp in zip(sm, keep)]
    dm = [m - m.mean(axis=0, keepdims=True) for m in dm]
    u, d, vt = np.linalg.svd(dm[0] @ dm[1].T, full_matrices=False)
    u, d, v = u[:, :ncomps], d[:ncomps], vt[:ncomps].T
    ul = dm[0].T @ u
    vl = dm[1].T @ v
    ul = ul / np.sqrt((ul * ul).sum(axis=0, keepdims=True))
    vl = vl / np.sqrt((vl * vl).sum(axis=0, keepdims=True))
    v0 = [np.var(m, axis=0, ddof=1).sum() for m in dm]
    nv = [np.var(dm[0] @ ul, axis=0, ddof=1) / v0[0], np.var(dm[1] @ vl, axis=0, ddof=1) / v0[1]]
    cw = np.sqrt(d)
    cw = cw / cw.max()
    return {"d": d, "u": u * cw[None, :], "v": v * cw[None, :], "ul": ul, "vl": vl, "nv": nv, "keep": keep,
            "genes": od}


def project_pair(samples: Sequence[Sample], od_genes: Sequence[str], rot: np.ndarray, var_scale: bool = True,
                 common_centering: bool = True):
    """R/conos.R:781-801.  With common.centering=TRUE the reference subtracts ONE NUMBER per sample
    (m[1] for the first sample, m[2] for the second -- `centering` is an atomic vector indexed with [[n]],
    SURVEY.md Appendix A 'centering quirk'); restated as written."""
    cproj = scaled_matrices(samples, od_genes, var_scale)
    if common_centering:
        ncells = np.array([x.shape[0] for x in cproj], dtype=np.float64)
        means = np.stack([np.asarray(x.mean(axis=0)).ravel() for x in cproj], axis=0)
        m = (means * ncells[:, None]).sum(axis=0) / ncells.sum()
        centering = [m[0], m[1]]
    else:
        centering = [np.asarray(x.mean(axis=0)).ravel() for x in cproj]
    out = []
    for x, c in zip(cproj, centering):
        out.append((x.toarray() - c) @ rot)
    return out


# ----------------------------------------------------------------------------------------------------
# getNeighborMatrix -- R/conos.R:848-886
# ----------------------------------------------------------------------------------------------------
def neighbor_matrix(p1: np.ndarray, p2: np.ndarray, k: int, k1: Optional[int] = None, matching: str = "mNN",
                    metric: str = "angular", l2_sigma: float = 1e5, cor_base: float = 1.0,
                    min_similarity: float = 1e-5):
    """Returns COO (i, j, x) of the n1 x n2 similarity matrix in TsparseMatrix order (= CSC order: by column j,
    then row i ascending [3P])."""
    if k1 is None:
        k1 = k
    n1, n2 = p1.shape[0], p2.shape[0]
    i12, d12 = cross_knn(p1, p2, k1, metric)  # per cell of sample 1: neighbours in sample 2
    i21, d21 = cross_knn(p2, p1, k1, metric)  # per cell of sample 2: neighbours in sample 1
    s12 = convert_distance_to_similarity(d12, metric, l2_sigma, cor_base)
    s21 = convert_distance_to_similarity(d21, metric, l2_sigma, cor_base)
    q1 = np.repeat(np.arange(n1), k1)
    v = i12.ravel() >= 0
    n12t = sp.csc_matrix((s12.ravel()[v], (q1[v], i12.ravel()[v])), shape=(n1, n2))  # t(n12): (query in 1, nbr in 2)
    q2 = np.repeat(np.arange(n2), k1)
    v = i21.ravel() >= 0
    n21 = sp.csc_matrix((s21.ravel()[v], (i21.ravel()[v], q2[v])), shape=(n1, n2))   # (nbr in 1, query in 2)
    if matching == "NN":
        adj = (n21 + n12t).tocsc()
        adj.data = adj.data / 2.0
    elif matching == "mNN":
        adj = n21.multiply(n12t).tocsc()
        adj.data = np.sqrt(adj.data)
    else:
        raise ValueError("Unrecognized type of NN matching:" + matching)
    adj.data[adj.data < min_similarity] = 0.0
    adj.eliminate_zeros()
    adj.sort_indices()
    if k1 > k:
        adj = reduce_edges_in_graph_iteratively(adj, k)
    adj = adj.tocsc()
    adj.eliminate_zeros()
    adj.sort_indices()
    coo = adj.tocoo()  # csc -> coo keeps column-major order
    return coo.row.astype(np.int32), coo.col.astype(np.int32), coo.data.astype(np.float64)


def pare_down_hub_edges(m: sp.csc_matrix, row_n: np.ndarray, k: int, klow: int = -1) -> np.ndarray:
    """src/edgeFilter.cpp:23-60 on a CSC matrix; returns the new @x, mutates row_n."""
    m = m.tocsc()
    x = np.ascontiguousarray(m.data, dtype=np.float64).copy()
    p = np.ascontiguousarray(m.indptr, dtype=np.int32)
    i = np.ascontiguousarray(m.indices, dtype=np.int32)
    lib().oracle_pare_down_hub_edges(m.shape[1], _p(p, ctypes.c_int), _p(i, ctypes.c_int), _p(x, ctypes.c_double),
                                     _p(row_n, ctypes.c_int), k, klow)
    return x


def _reduce_side(adj: sp.csc_matrix, k: int, klow: int) -> sp.csc_matrix:
    # adj.mtx[, order(diff(p), decreasing=TRUE)] -- R's order() is stable; then pare; then drop0
    nnz_col = np.diff(adj.indptr)
    perm = np.argsort(-nnz_col, kind="stable")
    a = adj[:, perm].tocsc()
    a.sort_indices()
    row_n = np.bincount(a.indices, minlength=a.shape[0]).astype(np.int32)  # tabulate(i+1)
    a.data = pare_down_hub_edges(a, row_n, k, klow)
    a.eliminate_zeros()
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    return a[:, inv].tocsc()


def reduce_edges_in_graph(adj: sp.csc_matrix, k: int, klow: Optional[int] = None) -> sp.csc_matrix:
    """R/conos.R:890-901.  The reference permutes, pares, transposes, permutes, pares, transposes back and
    restores names; un-permuting after each side is equivalent because paring only zeroes entries."""
    if klow is None:
        klow = k
    a = _reduce_side(adj.tocsc(), k, klow)
    at = _reduce_side(a.T.tocsc(), k, klow)
    out = at.T.tocsc()
    out.sort_indices()
    return out


def reduce_edges_in_graph_iteratively(adj: sp.csc_matrix, k: int, max_kdiff: int = 5, n_steps: int = 3):
    """R/conos.R:906-933."""
    adj = adj.tocsc()
    cc = np.diff(adj.indptr)
    rc = np.bincount(adj.indices, minlength=adj.shape[0])
    maxd = int(max(cc.max(initial=0), rc.max(initial=0)))
    if maxd <= k:
        return adj
    klow = max(min(k, 3), k - max_kdiff)
    n_steps = min(n_steps, r_round(maxd / k))
    if n_steps > 1:
        ks = np.exp(np.linspace(math.log(maxd), math.log(k), n_steps + 1))[1:]
        ks = [r_round(float(t)) for t in ks]
    else:
        ks = [k]
    for ki in ks:
        adj = reduce_edges_in_graph(adj, ki)
    cc = np.diff(adj.indptr)
    rc = np.bincount(adj.indices, minlength=adj.shape[0])
    maxd = int(max(cc.max(initial=0), rc.max(initial=0)))
    if maxd - k > max_kdiff:
        adj = reduce_edges_in_graph(adj, k, klow=klow)
    return adj


# ----------------------------------------------------------------------------------------------------
# local edges -- R/conos.R:587-616
# ----------------------------------------------------------------------------------------------------
def local_edges(pca: np.ndarray, k_self: int, k_self_weight: float, metric: str = "angular", l2_sigma: float = 1e5):
    """Returns (neighbour, query, weight) in TsparseMatrix order of xk (by query column, neighbour ascending).
    diag(xk) <- 0 ; drop0 removes the self entry AND any exact-zero distance."""
    idx, dist = self_knn(pca, k_self + 1, metric)
    n = pca.shape[0]
    q = np.repeat(np.arange(n), k_self + 1)
    nb = idx.ravel()
    dd = dist.ravel().astype(np.float64)
    keep = (nb >= 0) & (nb != q) & (dd != 0.0)
    nb, q, dd = nb[keep], q[keep], dd[keep]
    order = np.lexsort((nb, q))
    nb, q, dd = nb[order], q[order], dd[order]
    w = convert_distance_to_similarity(dd, metric, l2_sigma) * k_self_weight
    return nb.astype(np.int32), q.astype(np.int32), w


# ----------------------------------------------------------------------------------------------------
# graph assembly -- R/conclass.R:321-343  [3P: igraph graph_from_edgelist / simplify]
# ----------------------------------------------------------------------------------------------------
def assemble_graph(frm: np.ndarray, to: np.ndarray, w: np.ndarray, typ: np.ndarray):
    """el <- el[w > 0,]; graph_from_edgelist(character matrix): vertices numbered by first appearance
    scanning the table row by row (from_1, to_1, from_2, ...); simplify(weight='sum', type='first'):
    undirected multi-edges merged, loops removed, edges returned sorted by (low id, high id).
    Inputs are global cell indices; returns dict(vertices, from, to, weight, type) with from<to vertex ids."""
    keep = w > 0
    frm, to, w, typ = frm[keep], to[keep], w[keep], typ[keep]
    seq = np.empty(2 * len(frm), dtype=np.int64)
    seq[0::2] = frm
    seq[1::2] = to
    uniq, first = np.unique(seq, return_index=True)
    order = np.argsort(first, kind="stable")
    vertices = uniq[order]
    vid = np.empty(int(seq.max()) + 1 if len(seq) else 0, dtype=np.int64)
    vid[vertices] = np.arange(len(vertices))
    a, b = vid[frm], vid[to]
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    nl = lo != hi
    lo, hi, w, typ = lo[nl], hi[nl], w[nl], typ[nl]
    key = lo * max(len(vertices), 1) + hi
    o = np.argsort(key, kind="stable")
    key, lo, hi, w, typ = key[o], lo[o], hi[o], w[o], typ[o]
    start = np.ones(len(key), dtype=bool)
    start[1:] = key[1:] != key[:-1]
    gid = np.cumsum(start) - 1
    wsum = np.zeros(int(start.sum()))
    np.add.at(wsum, gid, w)  # sequential accumulation in original edge order
    return {"vertices": vertices, "from": lo[start], "to": hi[start], "weight": wsum, "type": typ[start]}


def adjacency_csr(g) -> sp.csr_matrix:
    """as_adj(graph, attr='weight'): symmetric matrix in vertex order."""
    n = len(g["vertices"])
    a = sp.coo_matrix((g["weight"], (g["from"], g["to"])), shape=(n, n))
    return (a + a.T).tocsr()


# ----------------------------------------------------------------------------------------------------
# edge re-balancing -- R/conos.R:936-967, src/edge_rebalancing.cpp
# ----------------------------------------------------------------------------------------------------
def sum_weight_matrix(weights, row_inds, col_inds, factor_levels, normalize=True):
    weights = np.ascontiguousarray(weights, dtype=np.float64)
    row_inds = np.ascontiguousarray(row_inds, dtype=np.int32)
    col_inds = np.ascontiguousarray(col_inds, dtype=np.int32)
    fl = np.ascontiguousarray(factor_levels, dtype=np.int32)
    n_cells, n_levels = len(fl), int(fl.max())
    m = np.zeros((n_cells, n_levels), order="F")
    lib().oracle_sum_weight_matrix(_p(weights, ctypes.c_double), _p(row_inds, ctypes.c_int), _p(col_inds, ctypes.c_int),
                                   len(weights), _p(fl, ctypes.c_int), n_cells, n_levels, int(normalize),
                                   _p(m, ctypes.c_double))
    return m


def adjust_weights_c(weights, row_inds, col_inds, factor_levels, dividers):
    w = np.ascontiguousarray(weights, dtype=np.float64).copy()
    row_inds = np.ascontiguousarray(row_inds, dtype=np.int32)
    col_inds = np.ascontiguousarray(col_inds, dtype=np.int32)
    fl = np.ascontiguousarray(factor_levels, dtype=np.int32)
    dv = np.asfortranarray(dividers, dtype=np.float64)
    lib().oracle_adjust_weights(_p(w, ctypes.c_double), _p(row_inds, ctypes.c_int), _p(col_inds, ctypes.c_int), len(w),
                                _p(fl, ctypes.c_int), len(fl), _p(dv, ctypes.c_double))
    return w


def adjust_weights_by_cell_balancing(adj: sp.spmatrix, factor_per_cell: np.ndarray, balance_weights: bool,
                                     same_factor_downweight: float = 1.0, n_iters: int = 50):
    """R/conos.R:936-967 on the symmetric adjacency (TsparseMatrix triplets in CSC order)."""
    coo = adj.tocsc().tocoo()
    ri, ci, w = coo.row.astype(np.int32), coo.col.astype(np.int32), coo.data.astype(np.float64)
    lv, fl = np.unique(np.asarray(factor_per_cell), return_inverse=True)  # as.factor + droplevels
    fl = (fl + 1).astype(np.int32)
    if balance_weights:
        for _ in range(n_iters):
            frac = sum_weight_matrix(w, ri, ci, fl)
            div = frac * (frac > 1e-10).sum(axis=1, keepdims=True)
            w = adjust_weights_c(w, ri, ci, fl, div)
    if abs(same_factor_downweight - 1) > 1e-5:
        same = fl[ri] == fl[ci]
        w[same] *= same_factor_downweight
    return ri, ci, w


def reference_wij(i: np.ndarray, j: np.ndarray, d: np.ndarray, perplexity: float, n: int) -> sp.csc_matrix:
    """src/edgeweights.cpp:160-184 (duplicates summed by setFromTriplets)."""
    i = np.ascontiguousarray(i, dtype=np.int32)
    j = np.ascontiguousarray(j, dtype=np.int32)
    d = np.ascontiguousarray(d, dtype=np.float64)
    of = np.empty(2 * len(i) + 1, dtype=np.int32)
    ot = np.empty(2 * len(i) + 1, dtype=np.int32)
    ow = np.empty(2 * len(i) + 1, dtype=np.float64)
    ne = lib().oracle_reference_wij(_p(i, ctypes.c_int), _p(j, ctypes.c_int), _p(d, ctypes.c_double), len(i),
                                    ctypes.c_double(perplexity), _p(of, ctypes.c_int), _p(ot, ctypes.c_int),
                                    _p(ow, ctypes.c_double))
    m = sp.coo_matrix((ow[:ne], (of[:ne], ot[:ne])), shape=(n, n)).tocsc()
    return m


def build_wij_matrix(adj: sp.spmatrix, perplexity: float = 50) -> sp.csc_matrix:
    """R/largeVis.R:30-34 (CsparseMatrix method): is = column index per entry, x@x^2 (squared again in C)."""
    a = adj.tocsc()
    a.sort_indices()
    is_ = np.repeat(np.arange(a.shape[1]), np.diff(a.indptr))
    return reference_wij(is_, a.indices, a.data ** 2, perplexity, a.shape[0])


# ----------------------------------------------------------------------------------------------------
# buildGraph -- R/conclass.R:161-371 (north_star API subset: PCA | CPCA | CCA, mNN | NN, angular | L2)
# ----------------------------------------------------------------------------------------------------
def build_graph(samples: Dict[str, Sample], k: int = 15, k_self: int = 10, k_self_weight: float = 0.1,
                alignment_strength: Optional[float] = None, space: str = "PCA", matching_method: str = "mNN",
                metric: str = "angular", k1: Optional[int] = None, l2_sigma: float = 1e5, var_scale: bool = True,
                ncomps: int = 40, n_odgenes: int = 2000, common_centering: bool = True,
                score_component_variance: bool = False, balance_edge_weights: bool = False,
                same_factor_downweight: float = 1.0, rotations: Optional[dict] = None,
                balancing_factor_per_cell: Optional[dict] = None, balancing_factor_per_sample: Optional[dict] = None,
                k_same_factor: Optional[int] = None):
    """Returns dict(graph=assemble_graph(...), pairs={name: reduction}, edges=(from, to, w, type) pre-merge).
    `rotations` may supply per-pair reductions (keyed 'A.vs.B') so that the kNN stage can be compared with
    the same rot/u/v on both sides (hard part 3)."""
    if space not in ("CPCA", "PCA", "CCA"):
        raise ValueError("only the following spaces are currently supported: [CPCA PCA CCA]")
    if matching_method not in ("mNN", "NN"):
        raise ValueError("only the following matching methods are currently supported: ['mNN' 'NN']")
    metric_id(metric)
    names = list(samples)
    max_ncells = max(s.counts.shape[0] for s in samples.values())
    k1, a = resolve_k1(k, k1, alignment_strength, max_ncells)
    cor_base = cor_base_of(a)
    offs, o = {}, 0
    for n in names:
        offs[n] = o
        o += samples[n].counts.shape[0]
    pairs = {}
    ef, et, ew = [], [], []
    for ia in range(len(names)):          # combn(names, 2) column order
        for ib in range(ia + 1, len(names)):
            na, nb = names[ia], names[ib]
            pr = [samples[na], samples[nb]]
            key = f"{na}.vs.{nb}"
            red = rotations.get(key) if rotations is not None else None
            if red is None:
                if space == "PCA":
                    red = quick_plain_pca(pr, ncomps, n_odgenes, var_scale, score_component_variance)
                elif space == "CPCA":
                    red = quick_cpca(pr, ncomps, n_odgenes, var_scale, score_component_variance)
                else:
                    red = quick_cca(pr, ncomps, n_odgenes, var_scale)
            if red is None:
                continue
            pairs[key] = red
            if space in ("PCA", "CPCA"):
                rot = red["CPC"]
                if ncomps <= rot.shape[1]:
                    rot = rot[:, :ncomps]
                p1, p2 = project_pair(pr, red["genes"], rot, var_scale, common_centering)
                r1 = np.arange(p1.shape[0])
                r2 = np.arange(p2.shape[0])
            else:
                p1, p2 = red["u"], red["v"]
                r1 = np.nonzero(red["keep"][0])[0]
                r2 = np.nonzero(red["keep"][1])[0]
            # k.cur (R/conclass.R:230-233): pairs inside one balancing.factor.per.sample level use k.same.factor
            k_cur = k
            if balancing_factor_per_sample is not None and \
                    balancing_factor_per_sample[na] == balancing_factor_per_sample[nb]:
                k_cur = min(k if k_same_factor is None else k_same_factor, k1)
            i, j, x = neighbor_matrix(p1, p2, k_cur, k1, matching_method, metric, l2_sigma, cor_base)
            ef.append(r1[i] + offs[na])
            et.append(r2[j] + offs[nb])
            ew.append(x)
    n_inter = sum(len(t) for t in ef)
    typ = [np.ones(n_inter, dtype=np.int32)]
    if k_self > 0:
        for n in names:
            nb, q, w = local_edges(samples[n].pca, k_self, k_self_weight, metric, l2_sigma)
            ef.append(nb + offs[n])
            et.append(q + offs[n])
            ew.append(w)
            typ.append(np.zeros(len(w), dtype=np.int32))
    frm = np.concatenate(ef).astype(np.int64) if ef else np.zeros(0, np.int64)
    to = np.concatenate(et).astype(np.int64) if et else np.zeros(0, np.int64)
    w = np.concatenate(ew) if ew else np.zeros(0)
    typ = np.concatenate(typ)
    g = assemble_graph(frm, to, w, typ)
    out = {"graph": g, "pairs": pairs, "edges": (frm, to, w, typ), "k1": k1, "cor_base": cor_base, "offsets": offs}
    # R/conclass.R:346-368: factor.per.cell = the caller's, else balancing.factor.per.sample[dataset], else the dataset
    if balance_edge_weights or balancing_factor_per_cell is not None or balancing_factor_per_sample is not None:
        adj = adjacency_csr(g)
        if balancing_factor_per_cell is not None:
            cells = np.concatenate([np.asarray(samples[n].cells, dtype=object) for n in names])
            fac = np.array([str(balancing_factor_per_cell[c]) for c in cells], dtype=object)
        elif balancing_factor_per_sample is not None:
            fac = np.concatenate([np.full(samples[n].counts.shape[0], str(balancing_factor_per_sample[n]), dtype=object)
                                  for n in names])
        else:
            fac = np.concatenate([np.full(samples[n].counts.shape[0], str(n), dtype=object) for n in names])
        out["balanced"] = adjust_weights_by_cell_balancing(adj, fac[g["vertices"]].astype(str), balance_edge_weights,
                                                           same_factor_downweight)
    return out


# ----------------------------------------------------------------------------------------------------
# graph diffusion -- src/propagate_labels.cpp:65-222
# ----------------------------------------------------------------------------------------------------
def smooth_count_matrix(edge_from, edge_to, edge_weights, count_matrix, is_label_fixed=None, max_n_iters=10,
                        diffusion_fading=1.0, diffusion_fading_const=0.1, tol=1e-3, normalize=False, return_info=False):
    """smooth_count_matrix (:177-222) on 0-based vertex ids = rows of count_matrix."""
    m = np.array(count_matrix, dtype=np.float64, order="F", copy=True)
    ef = np.ascontiguousarray(edge_from, dtype=np.int32)
    et = np.ascontiguousarray(edge_to, dtype=np.int32)
    w = np.ascontiguousarray(edge_weights, dtype=np.float64)
    if m.shape[0] == 0 or m.shape[1] == 0 or len(w) == 0:
        return (m, 0, 1e10) if return_info else m
    fx = None if is_label_fixed is None or len(is_label_fixed) == 0 else np.ascontiguousarray(is_label_fixed, np.uint8)
    nrm = ctypes.c_double(0)
    f = lib().oracle_smooth_matrix
    f.restype = ctypes.c_int
    it = f(_p(ef, ctypes.c_int), _p(et, ctypes.c_int), _p(w, ctypes.c_double), ctypes.c_longlong(len(w)), m.shape[0],
           _p(m, ctypes.c_double), m.shape[1], None if fx is None else _p(fx, ctypes.c_ubyte), int(max_n_iters),
           ctypes.c_double(diffusion_fading), ctypes.c_double(diffusion_fading_const), ctypes.c_double(tol),
           int(bool(normalize)), ctypes.byref(nrm))
    if it < 0:
        raise ValueError("Negative edge length")
    return (m, it, nrm.value) if return_info else m


def propagate_labels(edge_from, edge_to, edge_weights, n_vertices, label_of_vertex, n_labels, max_n_iters=10,
                     diffusion_fading=10.0, diffusion_fading_const=0.5, tol=5e-3, fixed_initial_labels=False,
                     return_info=False):
    """propagate_labels (:67-118) on integer ids: one-hot rows, smooth_count_matrix_c with normalize = true, rows
    divided by their sums."""
    lab = np.asarray(label_of_vertex)
    probs = np.zeros((n_vertices, n_labels), dtype=np.float64, order="F")
    has = lab >= 0
    probs[np.flatnonzero(has), lab[has]] = 1.0
    fixed = (has & bool(fixed_initial_labels)).astype(np.uint8)
    ef = np.ascontiguousarray(edge_from, dtype=np.int32)
    et = np.ascontiguousarray(edge_to, dtype=np.int32)
    w = np.ascontiguousarray(edge_weights, dtype=np.float64)
    it, nv = 0, 1e10
    if n_vertices and n_labels:
        nrm = ctypes.c_double(0)
        f = lib().oracle_smooth_matrix
        f.restype = ctypes.c_int
        it = f(_p(ef, ctypes.c_int), _p(et, ctypes.c_int), _p(w, ctypes.c_double), ctypes.c_longlong(len(w)),
               n_vertices, _p(probs, ctypes.c_double), n_labels, _p(fixed, ctypes.c_ubyte), int(max_n_iters),
               ctypes.c_double(diffusion_fading), ctypes.c_double(diffusion_fading_const), ctypes.c_double(tol), 1,
               ctypes.byref(nrm))
        if it < 0:
            raise ValueError("Negative edge length")
        nv = nrm.value
    with np.errstate(invalid="ignore", divide="ignore"):
        s = np.zeros(n_vertices)
        for c in range(n_labels):          # Rcpp::sum of the row, left to right
            s = s + probs[:, c]
        probs = np.asfortranarray(probs / s[:, None])
    return (probs, it, nv) if return_info else probs
