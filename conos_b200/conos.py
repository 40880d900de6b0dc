"""Host-side mirror of the reference's R6 class ``Conos`` for the joint-graph construction path.

Same method names, argument meaning, defaults and error messages as R/conclass.R (``Conos$new``,
``addSamples``, ``buildGraph``; R's dotted argument names become underscores).  All arithmetic runs in
libconosgpu.so through the C ABI (include/conos_gpu.h); this file only does what the R layer does around
the native calls: argument checks, pair enumeration and caching (``self.pairs[space]["A.vs.B"]``,
R/conclass.R:1035-1107), gene-set selection by name (R/conos.R:107-117) and result packaging.
"""
from __future__ import annotations

import ctypes as C
import time
import warnings
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import CgPair, CgParams, CgSample, MATCHING, METRICS, SPACES, ptr

SUPPORTED_SPACES = ["CPCA", "JNMF", "genes", "PCA", "PMA", "CCA"]   # what the reference accepts
DEVICE_SPACES = ["CPCA", "PCA", "CCA"]                               # what the B200 path builds (north_star)


@dataclass
class Pagoda2:
    """The fields of a pagoda2 object that buildGraph reads (R/access_wrappers.R:9,34,65,85,157)."""
    counts: sp.csc_matrix                 # cells x genes, log-scale expression
    genes: List[str]                      # colnames(counts)
    cells: List[str]                      # rownames(counts)
    varinfo_v: Optional[np.ndarray] = None    # misc$varinfo$v  (aligned with genes)
    varinfo_qv: Optional[np.ndarray] = None   # misc$varinfo$qv
    pca: Optional[np.ndarray] = None          # reductions$PCA, cells x nPcs
    od_genes: List[str] = field(default_factory=list)  # getOdGenes(): most overdispersed first

    def getOdGenes(self, n_odgenes: Optional[int] = None) -> List[str]:
        return self.od_genes if n_odgenes is None else self.od_genes[:n_odgenes]


class PairEntry(dict):
    """One entry of self$pairs[[space]] (R/conclass.R:1090-1092).  `genes` (the pair's od-gene names, the rownames of
    CPC) is materialised from the integer ids on first access: an atlas run caches thousands of these."""

    def __init__(self, od_ids, namer, **kw):
        super().__init__(**kw)
        self._od_ids, self._namer = od_ids, namer

    def __missing__(self, key):
        if key == "genes":
            self["genes"] = self._namer(self._od_ids)
            return dict.__getitem__(self, "genes")
        raise KeyError(key)

    def __contains__(self, key):
        return key == "genes" or dict.__contains__(self, key)


class LazyNames:
    """V(g)$name: cell barcodes in vertex order, gathered from the per-sample lists on first use."""

    def __init__(self, cell_names, vertices):
        self._cells, self._vertices, self._arr = cell_names, np.array(vertices, copy=True), None

    def _get(self):
        if self._arr is None:
            self._arr = self._cells[self._vertices]
        return self._arr

    def __len__(self):
        return len(self._vertices)

    def __getitem__(self, i):
        return self._get()[i]

    def __iter__(self):
        return iter(self._get())

    def __array__(self, dtype=None, copy=None):
        return self._get() if dtype is None else self._get().astype(dtype)


@dataclass
class ConosGraph:
    """What ``self$graph`` (an igraph) carries for findCommunities / embedGraph (R/conclass.R:336-343)."""
    names: object               # V(g)$name (array of str), vertex order = first appearance in the edge table
    vertices: np.ndarray        # global cell index of every vertex
    edge_from: np.ndarray       # vertex ids, from < to, sorted by (from, to)
    edge_to: np.ndarray
    weight: np.ndarray          # E(g)$weight
    type: np.ndarray            # E(g)$type : 1 inter-sample, 0 intra-sample
    adj_p: np.ndarray           # as_adj(g, attr='weight'): symmetric CSR/CSC in vertex order
    adj_i: np.ndarray
    adj_x: np.ndarray

    def as_adj(self) -> sp.csr_matrix:
        n = len(self.vertices)
        return sp.csr_matrix((self.adj_x, self.adj_i, self.adj_p), shape=(n, n))

    def scanpy_matrices(self, cell_ids=None) -> dict:
        """The graph matrices saveConosForScanPy writes (R/integrations.R:170-180, 219-232):
        graph.conn = as_adjacency_matrix(graph, attr='weight')[cell.ids, cell.ids] as a dgCMatrix (slots x, i, p, Dim)
        and graph.dist with x <- 1 - x.  `cell_ids`: the row names of the joint count matrix (cells in sample order);
        default = the vertex order.  Every cell id must be a vertex of the graph (R fails on a missing name too)."""
        a = self.as_adj().tocsc()
        if cell_ids is not None:
            pos = {n: t for t, n in enumerate(np.asarray(self.names, dtype=object))}
            try:
                perm = np.fromiter((pos[c] for c in cell_ids), np.int64, len(cell_ids))
            except KeyError as e:
                raise KeyError(f"cell {e.args[0]!r} is not a vertex of the graph") from None
            a = a[perm][:, perm].tocsc()
        a.sort_indices()
        shape = np.array(a.shape, np.int32)
        conn = {"data": a.data.astype(np.float64), "indices": a.indices.astype(np.int32),
                "indptr": a.indptr.astype(np.int32), "shape": shape}
        dist = {"data": 1.0 - conn["data"], "indices": conn["indices"], "indptr": conn["indptr"], "shape": shape}
        return {"graph_connectivities": conn, "graph_distances": dist}

    def save_for_scanpy(self, path: str, cell_ids=None) -> None:
        """Write graph_connectivities/{data,shape,indices,indptr} and graph_distances/... into an HDF5 file like
        saveConosForScanPy (alignment.graph part, R/integrations.R:219-232).  Needs h5py, as the reference needs
        rhdf5 (R/integrations.R:112-114)."""
        if not path.endswith(".h5"):
            raise ValueError(f"File {path} must have the file extension *.h5")
        try:
            import h5py
        except ImportError:
            raise ImportError('The package "h5py" is required for save_for_scanpy(). Please install.') from None
        mats = self.scanpy_matrices(cell_ids)
        with h5py.File(path, "w") as f:
            for group, m in mats.items():
                g = f.create_group(group)
                for key in ("data", "shape", "indices", "indptr"):
                    g.create_dataset(key, data=m[key])

    def vcount(self) -> int:
        return len(self.vertices)

    def ecount(self) -> int:
        return len(self.edge_from)


def r_round(x: float) -> int:
    return int(round(x))  # IEC 60559 half-to-even, like R


class Conos:
    def __init__(self, x: Optional[Dict[str, Pagoda2]] = None, n_cores: int = 1, device: int = 0, ctx=None):
        self.samples: Dict[str, Pagoda2] = {}
        self.pairs: Dict[str, Dict[str, dict]] = {}
        self.graph: Optional[ConosGraph] = None
        self.n_cores = n_cores      # kept for API parity; the device path ignores it
        self._ctx = ctx
        self._device = device
        self._panel = None
        self._panel_names: List[str] = []
        self._panel_key = None
        self._panel_col: Dict[str, Optional[np.ndarray]] = {}
        self._vocab = None
        self._od_membership = {}
        self._od_pair_memo = {}
        self._cell_names = None
        self._pinned = {}
        self.timings: Dict[str, float] = {}
        if x is not None:
            self.addSamples(x)

    # ------------------------------------------------------------------ R/conclass.R:93-112
    def addSamples(self, x: Dict[str, Pagoda2], replace: bool = False):
        if not isinstance(x, dict) or len(x) == 0:
            raise ValueError("The samples must be provided as a named list")
        if len(self.samples) + (0 if not replace else 0) + len(x) < 2 and not self.samples:
            raise ValueError("The list of samples must contain at least two samples")
        if replace:
            self.samples = {}
        dup = [n for n in x if n in self.samples]
        if dup:
            raise ValueError("Some of the names in the provided sample list are identical to already existing samples")
        for n, s in x.items():
            if not isinstance(s, Pagoda2):
                raise ValueError("only Pagoda2 samples are supported by the B200 path")
            if s.counts.shape != (len(s.cells), len(s.genes)):
                raise ValueError(f"sample {n}: counts must be cells x genes")
            self.samples[n] = s
        self._vocab = None
        self._od_membership = {}
        self._od_pair_memo = {}
        self._cell_names = None
        self._drop_panel()

    def _drop_panel(self):
        if self._panel is not None:
            self.ctx.lib.cg_panel_free(self._panel)
            self._panel = None
            self._panel_key = None

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = _lib.default_context(self._device)
        return self._ctx

    # ------------------------------------------------------------------ device-resident panel
    def _ensure_panel(self, names: Sequence[str], n_odgenes: int = 2000, need_counts=None, need_pca=None):
        """Upload the samples.  Only the gene universe the path can touch travels: the union of the samples' top
        n.odgenes overdispersed genes (every pair's gene set is drawn from two of those lists, R/conos.R:107-117).
        Real pagoda2 objects carry 20-30 thousand genes of which a few thousand qualify; the synthetic panels use all
        of theirs, in which case the arrays are passed through untouched (no copy).
        `need_counts` / `need_pca` (sets of sample names, None = all): a multi-GPU rank uploads the counts only of the
        samples its pairs touch and the PCA only of the samples whose local neighbours it computes; the others are
        placeholders that just number their cells."""
        key = (tuple(names), int(n_odgenes), None if need_counts is None else frozenset(need_counts),
               None if need_pca is None else frozenset(need_pca))
        if self._panel is not None and self._panel_key == key:
            return
        self._drop_panel()
        arr, keep = self._sample_array(names, n_odgenes, need_counts, need_pca)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.cg_panel_create(self.ctx.h, arr, len(names), C.byref(h)))
        self._panel = h
        self._panel_key = key

    def _sample_array(self, names: Sequence[str], n_odgenes: int, need_counts=None, need_pca=None):
        """The cg_sample table of the panel (host pointers into the samples' own arrays; `keep` holds the references
        that must outlive the native call)."""
        self._ensure_vocab()
        allnames, per = self._vocab
        in_u = np.zeros(len(allnames), bool)
        for n in names:
            in_u[per[n][1][:n_odgenes]] = True
        arr = (CgSample * len(names))()
        keep = []
        self._panel_col = {}
        self._od_pair_memo = {}        # the memo holds columns of the uploaded matrices
        self._h2d_bytes = 0
        for t, n in enumerate(names):
            s = self.samples[n]
            arr[t].n_cells, arr[t].n_genes = s.counts.shape
            arr[t].pca_ld = s.counts.shape[0]
            want_pca = s.pca is not None and (need_pca is None or n in need_pca)
            if want_pca:
                pca = np.asfortranarray(s.pca, np.float64)
                keep.append(pca)
                arr[t].pca = ptr(pca, C.c_double)
                arr[t].n_pcs = pca.shape[1]
                self._h2d_bytes += pca.nbytes
            if need_counts is not None and n not in need_counts:
                self._panel_col[n] = None
                continue        # placeholder: p == NULL
            # the caller's CSC object itself when it is one: scipy caches the sorted-indices check on the object, so
            # the scan over all non-zeros happens once per sample, not once per upload
            m = s.counts if sp.isspmatrix_csc(s.counts) else sp.csc_matrix(s.counts)
            col_of = per[n][0]                                  # vocabulary id -> column of this sample
            ids_of_col = np.empty(m.shape[1], np.int64)
            present = np.flatnonzero(col_of >= 0)
            ids_of_col[col_of[present]] = present
            kept = np.flatnonzero(in_u[ids_of_col])             # columns that belong to the universe, ascending
            v, qv = s.varinfo_v, s.varinfo_qv
            if len(kept) < m.shape[1]:
                m = m[:, kept]
                v = None if v is None else np.asarray(v)[kept]
                qv = None if qv is None else np.asarray(qv)[kept]
                pc = np.full(len(ids_of_col), -1, np.int32)
                pc[kept] = np.arange(len(kept), dtype=np.int32)
                self._panel_col[n] = pc
            else:
                self._panel_col[n] = None
            if not m.has_sorted_indices:
                m.sort_indices()
            p = np.ascontiguousarray(m.indptr, np.int32)
            i = np.ascontiguousarray(m.indices, np.int32)
            x = np.ascontiguousarray(m.data, np.float64)
            v = None if v is None else np.ascontiguousarray(v, np.float64)
            qv = None if qv is None else np.ascontiguousarray(qv, np.float64)
            keep += [p, i, x, v, qv]
            arr[t].n_cells, arr[t].n_genes = m.shape
            arr[t].p, arr[t].i, arr[t].x = ptr(p, C.c_int32), ptr(i, C.c_int32), ptr(x, C.c_double)
            arr[t].var_v, arr[t].var_qv = ptr(v, C.c_double), ptr(qv, C.c_double)
            self._h2d_bytes += p.nbytes + i.nbytes + x.nbytes + (0 if v is None else v.nbytes + qv.nbytes)
        self._panel_names = list(names)
        return arr, keep

    def _uploaded_genes(self, name: str, n_odgenes: int) -> int:
        """Columns of sample `name` inside the od-gene universe of the panel (what _ensure_panel uploads), known to
        every rank whether or not it holds the sample."""
        key = (name, int(n_odgenes), tuple(self._panel_names))
        cache = self.__dict__.setdefault("_n_uploaded", {})
        if key not in cache:
            allnames, per = self._vocab
            in_u = np.zeros(len(allnames), bool)
            for n in self._panel_names:
                in_u[per[n][1][:n_odgenes]] = True
            col_of = per[name][0]
            cache[key] = int(np.count_nonzero(in_u[np.flatnonzero(col_of >= 0)]))
        return cache[key]

    # ------------------------------------------------------------------ R/conos.R:107-117
    def _ensure_vocab(self):
        """Gene names -> integer ids in sort() order (C locale), once per sample set."""
        if self._vocab is not None:
            return
        allnames = sorted(set(g for s in self.samples.values() for g in s.genes))
        gid = {g: t for t, g in enumerate(allnames)}
        per = {}
        for n, s in self.samples.items():
            ids = np.fromiter((gid[g] for g in s.genes), np.int64, len(s.genes))
            col_of = np.full(len(allnames), -1, np.int32)
            col_of[ids] = np.arange(len(ids), dtype=np.int32)
            od = np.fromiter((gid[g] for g in s.od_genes if g in gid), np.int64)
            per[n] = (col_of, od)
        self._vocab = (allnames, per)

    def _common_od_genes(self, na: str, nb: str, n_odgenes: int):
        """commonOverdispersedGenes for the pair + the column of every od gene in both samples:
        table() of the two top-n od lists, sorted by count (decreasing, stable => ties keep name order),
        restricted to genes present in both samples, first n.odgenes.
        Gene ids are ranks in sort() order, so "count 2 first, then count 1, names ascending inside a count" is two
        flatnonzero calls on the per-sample membership vectors of the top-n lists (cached per n.odgenes)."""
        # pure function of the samples and of the uploaded columns: memoised (the atlas asks 4950 times per call)
        memo = self.__dict__.setdefault("_od_pair_memo", {})
        hit = memo.get((na, nb, n_odgenes))
        if hit is not None and hit[3] is self._panel_col.get(na) and hit[4] is self._panel_col.get(nb):
            return hit[:3]
        self._ensure_vocab()
        allnames, per = self._vocab
        col_a, od_a = per[na]
        col_b, od_b = per[nb]
        memb = self._od_membership
        for n, od in ((na, od_a), (nb, od_b)):
            if (n, n_odgenes) not in memb:
                m = np.zeros(len(allnames), np.int8)
                m[od[:n_odgenes]] = 1          # an od list never repeats a gene
                memb[(n, n_odgenes)] = m
        cnt = memb[(na, n_odgenes)] + memb[(nb, n_odgenes)]
        ids = np.concatenate([np.flatnonzero(cnt == 2), np.flatnonzero(cnt == 1)])
        ids = ids[(col_a[ids] >= 0) & (col_b[ids] >= 0)]
        ids = ids[:min(len(ids), n_odgenes)]
        ga, gb = col_a[ids], col_b[ids]
        pca_, pcb_ = self._panel_col.get(na), self._panel_col.get(nb)   # columns of the uploaded (subset) matrices
        if pca_ is not None:
            ga = pca_[ga]
        if pcb_ is not None:
            gb = pcb_[gb]
        out = (ids, np.ascontiguousarray(ga, np.int32), np.ascontiguousarray(gb, np.int32))
        memo[(na, nb, n_odgenes)] = out + (pca_, pcb_)
        return out

    def _gene_names(self, ids) -> List[str]:
        allnames = self._vocab[0]
        return [allnames[t] for t in ids]

    # ------------------------------------------------------------------ R/conclass.R:161-371
    def buildGraph(self, k: int = 15, k_self: int = 10, k_self_weight: float = 0.1,
                   alignment_strength: Optional[float] = None, space: str = "PCA", matching_method: str = "mNN",
                   metric: str = "angular", k1: Optional[int] = None, data_type: str = "counts",
                   l2_sigma: float = 1e5, var_scale: bool = True, ncomps: int = 40, n_odgenes: int = 2000,
                   matching_mask=None, exclude_samples: Optional[Sequence[str]] = None,
                   common_centering: bool = True, verbose: bool = False, base_groups=None,
                   score_component_variance: bool = False, snn: bool = False, balance_edge_weights: bool = False,
                   balancing_factor_per_cell=None, same_factor_downweight: float = 1.0,
                   k_same_factor: Optional[int] = None, balancing_factor_per_sample=None,
                   pair_shard: Optional[Sequence[int]] = None, assemble: bool = True, fetch: bool = True,
                   world=None, graph_on: str = "root", n_devices: Optional[int] = None):
        """`world` = (rank, world_size, torch.distributed group or None): one process per GPU.  The pair list is
        partitioned by the library (cg_partition_pairs: LPT with sample locality), every rank uploads only the samples
        its pairs touch, the per-rank edge tables are gathered with NCCL inside the library (cg_edges_allgatherv) in
        the reference's row order, and the graph is assembled on rank 0 (`graph_on="root"`; `"all"`: on every rank).
        `fetch=False` keeps the assembled graph on the device (self.graph is a size-only stub).
        `n_devices` = N runs the whole call over N GPUs from THIS process (cg_job: one host thread per device inside
        the library, NCCL between them) -- the form an R session uses; the panel is uploaded by the job on every call."""
        t_total = time.perf_counter()
        if space not in SUPPORTED_SPACES:
            raise ValueError("only the following spaces are currently supported: [" + " ".join(SUPPORTED_SPACES) + "]")
        if space not in DEVICE_SPACES:
            raise NotImplementedError(f"space='{space}' is outside the B200 hot path (PCA, CPCA, CCA)")
        if matching_method not in MATCHING:
            raise ValueError("only the following matching methods are currently supported: ['mNN' 'NN']")
        if metric not in METRICS:
            raise ValueError("only the following distance metrics are currently supported: ['L2' 'angular']")
        if data_type != "counts":
            raise NotImplementedError("only data.type='counts' is on the B200 path")
        if base_groups is not None or snn:
            raise NotImplementedError("base.groups / snn are outside the B200 hot path (SURVEY.md 8f)")
        if k1 is None:
            k1 = k
        if k_same_factor is None:
            k_same_factor = k                       # R default: k.same.factor = k
        names = [n for n in self.samples if not (exclude_samples and n in exclude_samples)]
        if alignment_strength is not None:
            alignment_strength = min(max(alignment_strength, 0.0), 1.0)
            max_cells = max(self.samples[n].counts.shape[0] for n in self.samples)
            k1 = max(r_round(max_cells * alignment_strength ** 2), k)
        else:
            alignment_strength = 0.0
        if k1 < k:
            raise ValueError("k1 must be >= k")
        cor_base = 1.0 + min(1.0, alignment_strength * 10.0)

        # ---- updatePairs: pair list in combn order (or matching.mask), cached reductions
        pair_names = []
        for ia in range(len(names)):
            for ib in range(ia + 1, len(names)):
                if matching_mask is not None and not (matching_mask.get((names[ia], names[ib])) or
                                                      matching_mask.get((names[ib], names[ia]))):
                    continue
                pair_names.append((names[ia], names[ib]))
        if len(pair_names) < 1:
            raise ValueError("insufficient number of comparable pairs")
        index = {n: t for t, n in enumerate(names)}
        n_cells = np.array([self.samples[n].counts.shape[0] for n in names], np.float64)
        lib, h = self.ctx.lib, self.ctx.h

        # ---- who does what (one rank: everything)
        rank, world_size, group = world if world is not None else (0, 1, None)
        n_pairs = len(pair_names)
        pa = np.array([index[a] for a, _ in pair_names], np.int32)
        pb_ = np.array([index[b] for _, b in pair_names], np.int32)
        pair_owner = np.zeros(n_pairs, np.int32)
        sample_owner = np.zeros(len(names), np.int32)
        need_counts = need_pca = None
        if world_size > 1:
            if pair_shard is not None:
                raise ValueError("pair_shard and world are mutually exclusive")
            self.ctx.comm_init(rank, world_size, group)
            cost = n_cells[pa] * n_cells[pb_]
            # what a rank pays the first time it meets a sample (its Gram matrix and, end to end, its upload), in the
            # unit of `cost` (cell pairs): measured on configs[1], ~0.005 * cells * genes^2
            n_genes_s = np.array([min(self.samples[n].counts.shape[1], 4 * n_odgenes) for n in names], np.float64)
            scost = 0.005 * n_cells * n_genes_s ** 2
            self.ctx.check(lib.cg_partition_pairs(ptr(pa, C.c_int32), ptr(pb_, C.c_int32), ptr(cost, C.c_double),
                                                  n_pairs, ptr(scost, C.c_double), len(names), world_size,
                                                  ptr(pair_owner, C.c_int32)))
            if k_self > 0:   # the local-neighbour searches (R/conos.R:589) are units of their own (~0.4 cells^2 in
                # pair-cost units: one direction, no reduction / projection); they go to the ranks the pairs left light
                lcost = 0.4 * n_cells ** 2
                pload = np.bincount(pair_owner, weights=cost, minlength=world_size).astype(np.float64)
                self.ctx.check(lib.cg_partition_units(ptr(lcost, C.c_double), len(names), world_size,
                                                      ptr(pload, C.c_double), ptr(sample_owner, C.c_int32)))
            mine = np.flatnonzero(pair_owner == rank)
            need_counts = {names[t] for t in set(pa[mine].tolist()) | set(pb_[mine].tolist())}
            need_pca = {names[t] for t in np.flatnonzero(sample_owner == rank)} if k_self > 0 else set()
            my_pairs = mine.tolist()
        else:
            my_pairs = list(range(n_pairs)) if pair_shard is None else list(pair_shard)
        if n_devices is not None:
            if world_size > 1 or pair_shard is not None or not assemble:
                raise ValueError("n_devices excludes world / pair_shard / assemble=False")
            return self._build_graph_job(int(n_devices), names, pair_names, index, space, ncomps, n_odgenes, k, k1,
                                         k_self, k_self_weight, matching_method, metric, l2_sigma, cor_base, var_scale,
                                         common_centering, score_component_variance, balance_edge_weights,
                                         balancing_factor_per_cell, balancing_factor_per_sample,
                                         same_factor_downweight, k_same_factor, fetch, t_total)
        t_0 = time.perf_counter()
        self._ensure_panel(names, n_odgenes, need_counts, need_pca)
        self.timings["panel_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        cache = self.pairs.setdefault(space, {})
        if world_size > 1 and not all(f"{a}.vs.{b}" in cache or f"{b}.vs.{a}" in cache for a, b in pair_names):
            # every sample's Gram matrix is computed once in the job, by the least loaded of the ranks that hold the
            # sample, and broadcast (cg_panel_share_grams); same tables on every rank
            holders = [sorted(set(pair_owner[(pa == t) | (pb_ == t)].tolist())) for t in range(len(names))]
            gram_owner = np.full(len(names), -1, np.int32)
            load = np.zeros(world_size)
            for t in np.argsort(-n_cells, kind="stable"):
                if holders[t]:
                    r = min(holders[t], key=lambda q: (load[q], q))
                    gram_owner[t] = r
                    load[r] += n_cells[t]
            ng = np.array([self._uploaded_genes(n, n_odgenes) for n in names], np.int32)
            self.ctx.check(lib.cg_panel_share_grams(h, self._panel, ptr(gram_owner, C.c_int32), ptr(ng, C.c_int32)))
            self.timings["share_grams_ms"] = 1e3 * (time.perf_counter() - t_0)

        # balancing.factor.per.sample: pairs inside one level search k.cur = min(k.same.factor, k1) neighbours
        # (R/conclass.R:230-233)
        if balancing_factor_per_sample is not None:
            missing = [n for n in names if n not in balancing_factor_per_sample]
            if missing:
                raise ValueError("balancing.factor.per.sample must name every sample")

        carr, keep_alive, used_pairs = self._pair_table(my_pairs, pair_names, index, cache, space, ncomps, n_odgenes,
                                                        k1, k_same_factor, balancing_factor_per_sample)

        P = self._params(k, k1, k_self, ncomps, space, matching_method, metric, var_scale, common_centering,
                         score_component_variance, k_self_weight, l2_sigma, cor_base)
        return self._build_graph_rest(P, carr, keep_alive, used_pairs, cache, names, n_pairs, pair_owner, sample_owner,
                                      rank, world_size, graph_on, space, score_component_variance, k_self, assemble,
                                      fetch, balance_edge_weights, balancing_factor_per_cell,
                                      balancing_factor_per_sample, same_factor_downweight, t_0, t_total)

    def _build_graph_job(self, n_devices, names, pair_names, index, space, ncomps, n_odgenes, k, k1, k_self,
                         k_self_weight, matching_method, metric, l2_sigma, cor_base, var_scale, common_centering,
                         score_component_variance, balance_edge_weights, balancing_factor_per_cell,
                         balancing_factor_per_sample, same_factor_downweight, k_same_factor, fetch, t_total):
        """The whole buildGraph as ONE native call over n_devices GPUs of this process (cg_job_build_graph)."""
        lib = self.ctx.lib if self._ctx is not None else _lib.load()
        job = self.__dict__.get("_job")
        if job is None or job.n_devices != n_devices or not job.h:
            job = _lib.Job(n_devices)
            self._job = job
        t_0 = time.perf_counter()
        self._panel_col = {}
        self._od_pair_memo = {}        # the memo holds columns of the uploaded matrices
        arr, keep = self._sample_array(names, n_odgenes)
        cache = self.pairs.setdefault(space, {})
        if balancing_factor_per_sample is not None:
            missing = [n for n in names if n not in balancing_factor_per_sample]
            if missing:
                raise ValueError("balancing.factor.per.sample must name every sample")
        carr, keep_alive, used_pairs = self._pair_table(list(range(len(pair_names))), pair_names, index, cache, space,
                                                        ncomps, n_odgenes, k1, k_same_factor,
                                                        balancing_factor_per_sample)
        P = CgParams()
        lib.cg_params_default(C.byref(P))
        P.k, P.k1, P.k_self, P.ncomps = k, k1, k_self, ncomps
        P.space, P.matching, P.metric = SPACES[space], MATCHING[matching_method], METRICS[metric]
        P.var_scale, P.common_centering = int(var_scale), int(common_centering)
        P.score_component_variance = int(score_component_variance)
        P.k_self_weight, P.l2_sigma, P.cor_base = k_self_weight, l2_sigma, cor_base
        self.timings["host_pairs_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        gh = C.c_void_p()
        if lib.cg_job_build_graph(job.h, arr, len(names), carr, len(used_pairs), C.byref(P), C.byref(gh)) != 0:
            raise _lib.ConosGpuError((lib.cg_job_last_error(job.h) or b"cg_job_build_graph failed").decode())
        self.timings["job_build_graph_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()

        def locate(t):
            dev, slot = C.c_int32(0), C.c_int32(0)
            lib.cg_job_pair_location(job.h, t, C.byref(dev), C.byref(slot))
            return job.contexts[dev.value], slot.value
        self._fill_cache(cache, used_pairs, space, score_component_variance, locate)
        self.timings["cache_fill_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        balance = None
        if balance_edge_weights or balancing_factor_per_cell is not None or balancing_factor_per_sample is not None:
            balance = dict(per_cell=balancing_factor_per_cell, per_sample=balancing_factor_per_sample,
                           balance_weights=bool(balance_edge_weights), downweight=float(same_factor_downweight))
        g = self._graph_out(gh, names, fetch, balance, job.contexts[0])
        self.timings["assemble_fetch_ms"] = 1e3 * (time.perf_counter() - t_0)
        self.graph = g
        self.timings["total_ms"] = 1e3 * (time.perf_counter() - t_total)
        del keep, keep_alive
        return g

    def _pair_table(self, my_pairs, pair_names, index, cache, space, ncomps, n_odgenes, k1, k_same_factor,
                    balancing_factor_per_sample):
        """cg_pair table of the listed pairs: gene columns, cached reductions (the cache-hit path of updatePairs,
        R/conclass.R:1062-1071) and k.cur (R/conclass.R:230-233)."""
        carr = (CgPair * max(len(my_pairs), 1))()
        keep_alive = []
        used_pairs = []
        for pi in my_pairs:
            na, nb = pair_names[pi]
            key = f"{na}.vs.{nb}"
            red = cache.get(key) or cache.get(f"{nb}.vs.{na}")
            if red is not None and space in ("PCA", "CPCA") and red.get("CPC") is not None:
                # od.genes <- rownames(cached CPC) (R/conclass.R:240-243): the cached entry fixes the gene set
                od, ga, gb = self._cached_od_genes(red, na, nb)
            else:
                od, ga, gb = self._common_od_genes(na, nb, n_odgenes)
            if len(od) < 5:
                continue  # quick* return NULL: the pair vanishes from the graph (R/conos.R:206,278,330)
            cp = carr[len(used_pairs)]
            cp.a, cp.b = index[na], index[nb]
            cp.n_genes = len(od)
            cp.genes_a, cp.genes_b = ptr(ga, C.c_int32), ptr(gb, C.c_int32)
            if balancing_factor_per_sample is not None and \
                    balancing_factor_per_sample[na] == balancing_factor_per_sample[nb]:
                cp.k = min(int(k_same_factor), k1)
            if red is not None and space in ("PCA", "CPCA") and red.get("CPC") is not None:
                rot = np.asfortranarray(red["CPC"], np.float64)
                if ncomps > rot.shape[1]:
                    warnings.warn(f"specified ncomps ({ncomps}) is greater than the cached version ({rot.shape[1]})")
                cp.rot, cp.rot_rows, cp.rot_cols = ptr(rot, C.c_double), rot.shape[0], rot.shape[1]
                keep_alive.append(rot)
            elif red is not None and space == "CCA" and red.get("u") is not None:
                # every cached column is used, whatever ncomps says now (R/conclass.R:268-270)
                u = np.asfortranarray(red["u"], np.float64)
                v = np.asfortranarray(red["v"], np.float64)
                if u.shape[1] != v.shape[1]:
                    raise ValueError(f"cached CCA of {key}: u and v differ in their number of columns")
                ru = np.ascontiguousarray(red["rows_u"], np.int32)
                rv = np.ascontiguousarray(red["rows_v"], np.int32)
                cp.u, cp.v = ptr(u, C.c_double), ptr(v, C.c_double)
                cp.rows_u, cp.rows_v = ptr(ru, C.c_int32), ptr(rv, C.c_int32)
                cp.n_u, cp.n_v, cp.uv_cols = u.shape[0], v.shape[0], u.shape[1]
                keep_alive += [u, v, ru, rv]
            keep_alive += [ga, gb]
            used_pairs.append((pi, key, od))
        return carr, keep_alive, used_pairs

    def _params(self, k, k1, k_self, ncomps, space, matching_method, metric, var_scale, common_centering,
                score_component_variance, k_self_weight, l2_sigma, cor_base):
        P = CgParams()
        self.ctx.lib.cg_params_default(C.byref(P))
        P.k, P.k1, P.k_self, P.ncomps = k, k1, k_self, ncomps
        P.space, P.matching, P.metric = SPACES[space], MATCHING[matching_method], METRICS[metric]
        P.var_scale, P.common_centering = int(var_scale), int(common_centering)
        P.score_component_variance = int(score_component_variance)
        P.k_self_weight, P.l2_sigma, P.cor_base = k_self_weight, l2_sigma, cor_base
        return P

    def _build_graph_rest(self, P, carr, keep_alive, used_pairs, cache, names, n_pairs, pair_owner, sample_owner, rank,
                          world_size, graph_on, space, score_component_variance, k_self, assemble, fetch,
                          balance_edge_weights, balancing_factor_per_cell, balancing_factor_per_sample,
                          same_factor_downweight, t_0, t_total):
        lib, h = self.ctx.lib, self.ctx.h
        edges = C.c_void_p()
        counts = np.zeros(max(len(used_pairs), 1), np.int64)
        self.timings["host_pairs_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        self.ctx.check(lib.cg_pairs_edges(h, self._panel, carr, len(used_pairs), C.byref(P), C.byref(edges),
                                          ptr(counts, C.c_int64)))
        self.timings["pairs_edges_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        self._fill_cache(cache, used_pairs, space, score_component_variance)
        self._last_pair_counts = {used_pairs[t][1]: int(counts[t]) for t in range(len(used_pairs))}
        self._last_edges = edges
        self._last_params = P
        self._last_names = names
        self.timings["cache_fill_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        if not assemble:
            return edges

        # ---- local edges (R/conclass.R:328-333) and the exchange (R: rbind of the per-pair data frames, :316-321)
        n_inter = int(sum(self._last_pair_counts.values()))
        if world_size > 1:
            # one allgatherv for everything: segments = the pairs (combn order), then the samples' local edges
            n_s = len(names)
            seg_owner = np.concatenate([pair_owner, sample_owner]).astype(np.int32)
            seg_counts = np.zeros(n_pairs + n_s, np.int64)
            for t, (pi, _, _) in enumerate(used_pairs):
                seg_counts[pi] = counts[t]
            if k_self > 0:
                sidx = np.flatnonzero(sample_owner == rank).astype(np.int32)
                mine_counts = np.zeros(max(len(sidx), 1), np.int64)
                self.ctx.check(lib.cg_local_edges(h, self._panel, ptr(sidx, C.c_int32), len(sidx), C.byref(P),
                                                  C.byref(edges), ptr(mine_counts, C.c_int64)))
                seg_counts[n_pairs + sidx] = mine_counts[:len(sidx)]
            self.ctx.check(lib.cg_edges_allgatherv(h, C.byref(edges), ptr(seg_owner, C.c_int32),
                                                   ptr(seg_counts, C.c_int64), n_pairs + n_s))
            n_inter = self._count_type1(edges) if k_self > 0 else int(lib.cg_edges_count(edges))
        elif k_self > 0:
            sidx = np.arange(len(names), dtype=np.int32)
            self.ctx.check(lib.cg_local_edges(h, self._panel, ptr(sidx, C.c_int32), len(names), C.byref(P),
                                              C.byref(edges), None))
        self._n_inter_edges = n_inter
        self.timings["local_edges_ms"] = 1e3 * (time.perf_counter() - t_0)
        t_0 = time.perf_counter()
        if world_size > 1 and graph_on == "root" and rank != 0:
            lib.cg_edges_free(edges)
            self._last_edges = None
            self.graph = None
            self.timings["assemble_fetch_ms"] = 0.0
            self.timings["total_ms"] = 1e3 * (time.perf_counter() - t_total)
            return None
        balance = None
        if balancing_factor_per_sample is not None and balancing_factor_per_cell is not None:
            warnings.warn("Both balancing.factor.per.cell and balancing.factor.per.sample are provided. "
                          "Used the former for balancing edge weights")
        if balance_edge_weights or balancing_factor_per_cell is not None or balancing_factor_per_sample is not None:
            balance = dict(per_cell=balancing_factor_per_cell, per_sample=balancing_factor_per_sample,
                           balance_weights=bool(balance_edge_weights), downweight=float(same_factor_downweight))
        g = self._assemble(edges, names, fetch=fetch, balance=balance)
        self.timings["assemble_fetch_ms"] = 1e3 * (time.perf_counter() - t_0)
        lib.cg_edges_free(edges)
        self._last_edges = None
        self.graph = g
        self.timings["total_ms"] = 1e3 * (time.perf_counter() - t_total)
        return g

    def _count_type1(self, edges) -> int:
        """Inter-sample rows of a gathered edge table = the rows before the first local edge: the local segments follow
        the pairs' segments, and the library knows the total, so the count comes from the segment totals it returned."""
        return int(self.ctx.lib.cg_edges_count_type(self.ctx.h, edges, 1))

    def _cached_od_genes(self, red, na, nb):
        """Gene set of a cached PCA / CPCA entry = rownames(CPC) (R/conclass.R:240-243), mapped to the columns of the
        two uploaded samples."""
        self._ensure_vocab()
        allnames, per = self._vocab
        ids = getattr(red, "_od_ids", None)
        if ids is None:
            gid = {g: t for t, g in enumerate(allnames)}
            ids = np.array([gid[g] for g in red["genes"]], np.int64)
        ids = np.asarray(ids, np.int64)
        ga, gb = per[na][0][ids], per[nb][0][ids]
        if (ga < 0).any() or (gb < 0).any():
            raise ValueError(f"cached reduction of {na}.vs.{nb} names genes that a sample does not have")
        pca_, pcb_ = self._panel_col.get(na), self._panel_col.get(nb)
        if pca_ is not None:
            ga = pca_[ga]
        if pcb_ is not None:
            gb = pcb_[gb]
        if (np.asarray(ga) < 0).any() or (np.asarray(gb) < 0).any():
            raise ValueError(f"cached reduction of {na}.vs.{nb} uses genes outside the uploaded od-gene universe "
                             "(it was computed with a larger n.odgenes)")
        return ids, np.ascontiguousarray(ga, np.int32), np.ascontiguousarray(gb, np.int32)

    def _fill_cache(self, cache, used_pairs, space, score_component_variance, locate=None):
        """populate self$pairs[[space]] like updatePairs does (R/conclass.R:1090-1092): PCA list(CPC, nv);
        CPCA list(D, CPC, niter, nv) (src/cpca.cpp:93-100); CCA list(d, u, v, ul, vl, nv) (R/conos.R:344-366).
        `locate(t)` -> (context, slot) of the t-th used pair (a multi-device job keeps a reduction on its device)."""
        # PCA without variance scores on one context (the atlas case: thousands of pairs): all rotations in two calls
        if space == "PCA" and not score_component_variance and locate is None and len(used_pairs) > 8:
            cx = self.ctx
            n = len(used_pairs)
            nr, nc = np.zeros(n, np.int32), np.zeros(n, np.int32)
            off = np.zeros(n + 1, np.int64)
            cx.check(cx.lib.cg_pair_reductions_all(cx.h, ptr(nr, C.c_int32), ptr(nc, C.c_int32), ptr(off, C.c_int64),
                                                   None, 0))
            buf = np.empty(int(off[n]), np.float64)
            cx.check(cx.lib.cg_pair_reductions_all(cx.h, None, None, None, ptr(buf, C.c_double), buf.size))
            for t, (pi, key, od) in enumerate(used_pairs):
                if key not in cache:
                    rot = buf[off[t]:off[t + 1]].reshape((nr[t], nc[t]), order="F")
                    cache[key] = PairEntry(od, self._gene_names, CPC=rot)
            return
        for t, (pi, key, od) in enumerate(used_pairs):
            if key in cache:
                continue
            cx, slot = (self.ctx, t) if locate is None else locate(t)
            lib, h = cx.lib, cx.h
            if space == "CCA":
                nu, nvv, dd = C.c_int32(0), C.c_int32(0), C.c_int32(0)
                cx.check(lib.cg_pair_cca(h, slot, C.byref(nu), C.byref(nvv), C.byref(dd), None, None, None, None,
                                               None))
                if nu.value == 0:
                    continue
                u = np.zeros((nu.value, dd.value), order="F")
                v = np.zeros((nvv.value, dd.value), order="F")
                ru = np.zeros(nu.value, np.int32)
                rv = np.zeros(nvv.value, np.int32)
                sg = np.zeros(dd.value)
                cx.check(lib.cg_pair_cca(h, slot, None, None, None, ptr(u, C.c_double), ptr(v, C.c_double),
                                               ptr(ru, C.c_int32), ptr(rv, C.c_int32), ptr(sg, C.c_double)))
                cache[key] = PairEntry(od, self._gene_names, u=u, v=v, rows_u=ru, rows_v=rv, d=sg)
                ng, dl = C.c_int32(0), C.c_int32(0)
                cx.check(lib.cg_pair_cca_loadings(h, slot, C.byref(ng), C.byref(dl), None, None, None))
                if ng.value > 0:
                    ul = np.zeros((ng.value, dl.value), order="F")
                    vl = np.zeros((ng.value, dl.value), order="F")
                    nv = np.zeros((2, dl.value))
                    cx.check(lib.cg_pair_cca_loadings(h, slot, None, None, ptr(ul, C.c_double),
                                                            ptr(vl, C.c_double), ptr(nv, C.c_double)))
                    cache[key].update(ul=ul, vl=vl, nv=[nv[0], nv[1]])
                continue
            nr, nc = C.c_int32(0), C.c_int32(0)
            cx.check(lib.cg_pair_reduction(h, slot, C.byref(nr), C.byref(nc), None, None))
            rot = np.empty((nr.value, nc.value), order="F")
            nv = np.zeros((2, nc.value)) if score_component_variance else None
            cx.check(lib.cg_pair_reduction(h, slot, None, None, ptr(rot, C.c_double), ptr(nv, C.c_double)))
            cache[key] = PairEntry(od, self._gene_names, CPC=rot)
            if space == "CPCA":
                cache[key]["ncomp"] = nc.value
                D = np.zeros((nc.value, 2), order="F")   # res$D = t(D): ncomp x k (src/cpca.cpp:97)
                niter = np.zeros(nc.value)
                cx.check(lib.cg_pair_cpca(h, slot, None, ptr(D, C.c_double), ptr(niter, C.c_double)))
                cache[key].update(D=D, niter=niter)
            if nv is not None:
                rr = nc.value // 2 if space == "PCA" else nc.value
                cache[key]["nv"] = [nv[0, :rr], nv[1, :rr]]

    # ------------------------------------------------------------------ diffusion over the graph
    def propagateLabels(self, labels: Dict[str, object], method: str = "diffusion", max_iters: int = 100,
                        diffusion_fading: float = 10.0, diffusion_fading_const: float = 0.1, tol: float = 0.025,
                        fixed_initial_labels: bool = True) -> dict:
        """Conos$propagateLabels(labels, method='diffusion') (R/conclass.R:910-925) -> propagateLabelsDiffusion
        (R/conos.R:1071-1083) -> propagate_labels (src/propagate_labels.cpp:67-118) on the device.
        `labels`: cell name -> label.  Returns labels (per vertex, first maximum like which.max), uncertainty
        (1 - max probability), label_distribution (n_vertices x n_labels, rows = V(graph)$name order) and label_names
        (numbered by first appearance among the labelled cells that are in the graph, like order_strings)."""
        if method == "solver":
            raise NotImplementedError("method='solver' (a sparse Laplacian solve on the CPU) is outside the device path")
        if method != "diffusion":
            raise ValueError(f"Unknown method: {method}. Only 'solver' and 'diffusion' are supported.")
        if self.graph is None:
            raise ValueError("run buildGraph() first")
        from . import ops
        g = self.graph
        names = np.asarray(g.names, dtype=object)
        pos = {n: t for t, n in enumerate(names)}
        label_ids: Dict[object, int] = {}
        lab = np.full(len(names), -1, np.int32)
        for cell, l in labels.items():          # labels[intersect(names(labels), V(graph)$name)]: the caller's order
            t = pos.get(cell)
            if t is None:
                continue
            lab[t] = label_ids.setdefault(l, len(label_ids))
        dist = ops.propagate_labels(g.edge_from, g.edge_to, g.weight, len(names), lab, len(label_ids),
                                    max_n_iters=max_iters, diffusion_fading=diffusion_fading,
                                    diffusion_fading_const=diffusion_fading_const, tol=tol,
                                    fixed_initial_labels=fixed_initial_labels, ctx=self.ctx)
        label_names = list(label_ids)
        # which.max / max skip nothing in R: a row of NaN (no label reaches the vertex) gives integer(0) / NaN there;
        # here such vertices get label None and uncertainty NaN
        best = np.zeros(len(names), np.int64)
        conf = np.full(len(names), np.nan)
        if dist.shape[1] > 0:
            ok = ~np.isnan(dist).any(axis=1)
            best[ok] = np.argmax(dist[ok], axis=1)
            conf[ok] = dist[ok].max(axis=1)
        out_labels = np.array([label_names[b] if o else None for b, o in zip(best, ~np.isnan(conf))], dtype=object)
        return {"labels": out_labels, "uncertainty": 1.0 - conf, "label_distribution": dist,
                "label_names": label_names, "cells": names}

    def _od_genes_uniformly(self, n_genes: int) -> List[str]:
        """getOdGenesUniformly (R/conos.R:505-513): genes by their best rank in any sample's od list; ties in gene-name
        order (group_by sorts the groups, arrange is stable)."""
        rank: Dict[str, int] = {}
        for s in self.samples.values():
            for r, gname in enumerate(s.od_genes):
                if gname not in rank or r < rank[gname]:
                    rank[gname] = r
        genes = sorted(rank, key=lambda gname: (rank[gname], gname))
        return genes[:min(n_genes, len(genes))]

    def correctGenes(self, genes: Optional[Sequence[str]] = None, n_od_genes: int = 500, fading: float = 10.0,
                     fading_const: float = 0.5, max_iters: int = 15, tol: float = 5e-3, name: str = "diffusion",
                     count_matrix: Optional[np.ndarray] = None, count_matrix_cells: Optional[Sequence[str]] = None,
                     normalize: bool = True) -> np.ndarray:
        """Conos$correctGenes (R/conclass.R:862-893): expression smoothed over the joint graph by smooth_count_matrix
        (src/propagate_labels.cpp:177-222).  Default matrix: cells x genes of the samples' count matrices over the
        genes common to all samples and in `genes` (default: getOdGenesUniformly(n.od.genes)), rows reordered to
        V(graph)$name.  `count_matrix` (genes x cells like the reference's argument) needs `count_matrix_cells`.
        Note: like the reference this passes tol only through the default of smooth_count_matrix (1e-3) -- `tol` is an
        argument of correctGenes that the reference never forwards (R/conclass.R:890-891)."""
        if self.graph is None:
            raise ValueError("run buildGraph() first")
        from . import ops
        g = self.graph
        vn = np.asarray(g.names, dtype=object)
        if count_matrix is None:
            if genes is None:
                genes = self._od_genes_uniformly(n_od_genes)
            common = None
            for s in self.samples.values():
                gs = set(s.genes)
                common = gs if common is None else common & gs
            first = next(iter(self.samples.values()))
            keep = [gname for gname in first.genes if gname in common and gname in set(genes)]   # Reduce(intersect) order
            blocks, cells = [], []
            for s in self.samples.values():
                col = {gname: t for t, gname in enumerate(s.genes)}
                idx = np.array([col[gname] for gname in keep], dtype=np.int64)
                blocks.append(np.asarray(sp.csc_matrix(s.counts)[:, idx].todense(), dtype=np.float64))
                cells.extend(s.cells)
            cm = np.vstack(blocks)
            self._corrected_genes = keep
        else:
            if count_matrix_cells is None:
                raise ValueError("count_matrix needs count_matrix_cells (its column names)")
            cm = np.asarray(count_matrix, dtype=np.float64).T
            cells = list(count_matrix_cells)
        row = {c: t for t, c in enumerate(cells)}
        if not all(c in row for c in vn):
            raise ValueError("count.matrix does not provide values for all the vertices in the alignment graph!")
        cm = cm[np.array([row[c] for c in vn], dtype=np.int64)]
        res = ops.smooth_count_matrix(g.edge_from, g.edge_to, g.weight, cm, max_n_iters=max_iters,
                                      diffusion_fading=fading, diffusion_fading_const=fading_const, normalize=normalize,
                                      ctx=self.ctx)
        if not hasattr(self, "expression_adj"):
            self.expression_adj = {}
        self.expression_adj[name] = res
        return res

    def close(self):
        """Release the device-resident panel (R: the object going out of scope / rm(con); gc())."""
        try:
            if self._panel is not None and self._ctx is not None and getattr(self._ctx, "h", None):
                self._drop_panel()
        except Exception:
            pass
        self._panel = None

    def __del__(self):
        try:
            import sys
            if not sys.is_finalizing():
                self.close()
        except Exception:      # interpreter teardown: the import machinery may already be gone
            pass

    def _out(self, key: str, n: int, dtype):
        """Result buffer in pinned host memory when torch is importable (grow-only, reused across calls; the
        returned array is a private copy-free view valid until the next buildGraph), plain NumPy otherwise."""
        try:
            import torch
            if not torch.cuda.is_available():
                raise RuntimeError
            tdt = {np.int32: torch.int32, np.int64: torch.int64, np.float64: torch.float64}[dtype]
            buf = self._pinned.get(key)
            if buf is None or buf.numel() < n or buf.dtype != tdt:
                buf = torch.empty(max(n, 1), dtype=tdt).pin_memory()
                self._pinned[key] = buf
            return buf.numpy()[:n]
        except Exception:
            return np.empty(n, dtype)

    def _assemble(self, edges, names, fetch: bool = True, balance=None) -> ConosGraph:
        n_total = sum(self.samples[n].counts.shape[0] for n in names)
        gh = C.c_void_p()
        self.ctx.check(self.ctx.lib.cg_assemble_graph(self.ctx.h, edges, n_total, C.byref(gh)))
        return self._graph_out(gh, names, fetch, balance, self.ctx)

    def _graph_out(self, gh, names, fetch, balance, cx) -> ConosGraph:
        """Optional balancing, then the host copy of an assembled graph that lives on context `cx`."""
        lib, h = cx.lib, cx.h
        nv, ne = C.c_int32(0), C.c_int64(0)
        lib.cg_graph_sizes(gh, C.byref(nv), C.byref(ne))
        if balance is not None and nv.value > 0:
            self._balance_on_device(gh, nv.value, names, balance, cx)
        if not fetch:
            lib.cg_graph_free(gh)
            z = np.zeros(0, np.int32)
            g = ConosGraph(names=[], vertices=z, edge_from=z, edge_to=z, weight=np.zeros(0), type=z,
                           adj_p=np.zeros(1, np.int64), adj_i=z, adj_x=np.zeros(0))
            g.n_vertices_device, g.n_edges_device = nv.value, ne.value
            return g
        vertices = self._out("v", nv.value, np.int32)
        ef = self._out("ef", ne.value, np.int32)
        et = self._out("et", ne.value, np.int32)
        w = self._out("w", ne.value, np.float64)
        ty = self._out("ty", ne.value, np.int32)
        cx.check(lib.cg_graph_fetch(h, gh, ptr(vertices, C.c_int32), ptr(ef, C.c_int32), ptr(et, C.c_int32),
                                          ptr(w, C.c_double), ptr(ty, C.c_int32)))
        ap = self._out("ap", nv.value + 1, np.int64)
        ai = self._out("ai", 2 * ne.value, np.int32)
        ax = self._out("ax", 2 * ne.value, np.float64)
        cx.check(lib.cg_graph_fetch_adjacency(h, gh, ptr(ap, C.c_int64), ptr(ai, C.c_int32), ptr(ax, C.c_double)))
        lib.cg_graph_free(gh)
        if self._cell_names is None or self._cell_names[0] != list(names):
            self._cell_names = (list(names), np.array([c for n in names for c in self.samples[n].cells], dtype=object))
        return ConosGraph(names=LazyNames(self._cell_names[1], vertices), vertices=vertices, edge_from=ef, edge_to=et, weight=w,
                          type=ty, adj_p=ap, adj_i=ai, adj_x=ax)

    def _balance_on_device(self, gh, n_vertices, names, balance, cx=None):
        """adjustWeightsByCellBalancing (R/conclass.R:346-368, R/conos.R:936-967) as one device call on the assembled
        graph.  factor.per.cell: the caller's (cell name -> level), else balancing.factor.per.sample[dataset of the
        cell], else the dataset itself (getDatasetPerCell); levels are numbered like as.factor() %>% droplevels()
        (sorted unique values present in the graph)."""
        cx = cx or self.ctx
        lib, h = cx.lib, cx.h
        vertices = np.empty(n_vertices, np.int32)
        cx.check(lib.cg_graph_fetch(h, gh, ptr(vertices, C.c_int32), None, None, None, None))
        sizes = [self.samples[n].counts.shape[0] for n in names]
        sample_of = np.repeat(np.arange(len(names)), sizes)[vertices]
        if balance["per_cell"] is not None:
            if self._cell_names is None or self._cell_names[0] != list(names):
                self._cell_names = (list(names),
                                    np.array([c for n in names for c in self.samples[n].cells], dtype=object))
            cells = self._cell_names[1][vertices]
            try:
                values = [balance["per_cell"][c] for c in cells]
            except KeyError as e:
                raise ValueError(f"balancing.factor.per.cell has no entry for cell {e}") from None
        elif balance["per_sample"] is not None:
            values = [balance["per_sample"][names[t]] for t in sample_of]
        else:
            values = [names[t] for t in sample_of]
        levels, codes = np.unique(np.asarray([str(v) for v in values]), return_inverse=True)
        fac = np.ascontiguousarray(codes + 1, np.int32)
        cx.check(lib.cg_graph_balance(h, gh, ptr(fac, C.c_int32), len(levels), int(balance["balance_weights"]),
                                            balance["downweight"], 50))
