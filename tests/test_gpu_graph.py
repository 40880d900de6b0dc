"""GPU parity of matching, local edges, projection and graph assembly against the oracle (via the C ABI and the
Conos host mirror).  Integer/index results bit-exact; weights to 1e-12 relative (same float32 distances, double
arithmetic afterwards); projections to 1e-10 (fp64 sums in a different order)."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import to_oracle_samples

pytestmark = pytest.mark.gpu


def mixture(rng, n, d, k=8, spread=3.0):
    cent = rng.normal(0, spread, (k, d))
    lab = rng.integers(0, k, n)
    return cent[lab] + rng.normal(0, 1, (n, d))


@pytest.mark.parametrize("n1,n2,d,k,cor_base", [(25, 34, 30, 15, 1.0), (700, 500, 30, 15, 1.0), (400, 900, 50, 30, 2.0)])
def test_neighbor_matrix_mnn(ctx, oracle, n1, n2, d, k, cor_base):
    from conos_b200 import ops
    rng = np.random.default_rng(n1 + n2)
    p1, p2 = mixture(rng, n1, d), mixture(rng, n2, d)
    i, j, x = ops.getNeighborMatrix(p1, p2, k, metric="angular", cor_base=cor_base, ctx=ctx)
    oi, oj, ox = oracle.neighbor_matrix(p1, p2, k, metric="angular", cor_base=cor_base)
    assert np.array_equal(i, oi) and np.array_equal(j, oj)
    np.testing.assert_allclose(x, ox, rtol=1e-12, atol=0)
    assert len(i) > 0 and x.min() >= 1e-5


@pytest.mark.parametrize("matching,k,k1", [("mNN", 12, 12), ("NN", 8, 8), ("mNN", 6, 30)])
def test_neighbor_matrix_l2_metric(ctx, oracle, matching, k, k1):
    """metric = 'L2': squared Euclidean distances, similarity exp(-d / l2.sigma) (R/conos.R:767-773) -- matching,
    the min.similarity cut and hub paring on L2 similarities, against the oracle."""
    from conos_b200 import ops
    rng = np.random.default_rng(31 + k)
    p1, p2 = rng.normal(size=(450, 25)), rng.normal(size=(380, 25)) + 0.2
    for sigma in (1e5, 30.0):
        i, j, x = ops.getNeighborMatrix(p1, p2, k, k1, matching, "L2", l2_sigma=sigma, ctx=ctx)
        oi, oj, ox = oracle.neighbor_matrix(p1, p2, k, k1, matching, "L2", sigma)
        assert np.array_equal(i, oi) and np.array_equal(j, oj)
        np.testing.assert_allclose(x, ox, rtol=1e-12)


def test_build_graph_l2_metric(ctx, oracle, small_panel):
    """buildGraph(metric = 'L2') end to end with the oracle's rotations: inter-sample and local edges carry
    exp(-d / l2.sigma) similarities."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, metric="L2", l2_sigma=50.0)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.pairs["PCA"] = {key: {"CPC": red["CPC"], "genes": red["genes"]} for key, red in ref["pairs"].items()}
    g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, metric="L2", l2_sigma=50.0)
    _graphs_equal(g, ref["graph"], rtol=1e-6)
    con.close()


def test_neighbor_matrix_negative_correlation_dropped(ctx, oracle):
    """distance >= cor.base gives similarity 0: such pairs are dropped even when mutual (drop0)."""
    from conos_b200 import ops
    rng = np.random.default_rng(1)
    p1 = rng.normal(0, 1, (40, 5))
    p2 = -p1[:30] + rng.normal(0, 0.01, (30, 5))
    i, j, x = ops.getNeighborMatrix(p1, p2, 10, ctx=ctx)
    oi, oj, ox = oracle.neighbor_matrix(p1, p2, 10)
    assert np.array_equal(i, oi) and np.array_equal(j, oj)
    np.testing.assert_allclose(x, ox, rtol=1e-12)


def _graphs_equal(g, og, rtol=1e-12):
    assert np.array_equal(g.vertices, og["vertices"])
    assert np.array_equal(g.edge_from, og["from"]) and np.array_equal(g.edge_to, og["to"])
    assert np.array_equal(g.type, og["type"])
    np.testing.assert_allclose(g.weight, og["weight"], rtol=rtol, atol=0)


@pytest.mark.parametrize("common_centering,var_scale", [(True, True), (False, True), (True, False)])
def test_build_graph_with_oracle_rotations(ctx, oracle, small_panel, common_centering, var_scale):
    """Same rot on both sides (SURVEY hard part 3): the whole path after the reduction must agree."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300,
                             common_centering=common_centering, var_scale=var_scale)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.pairs["PCA"] = {key: {"CPC": red["CPC"], "genes": red["genes"]} for key, red in ref["pairs"].items()}
    ctx.lib.cg_keep_projections(ctx.h, 1)
    g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, common_centering=common_centering,
                       var_scale=var_scale)
    ctx.lib.cg_keep_projections(ctx.h, 0)
    # projections of the last pair, against the oracle's dense restatement
    names = list(small_panel)
    pr = [osamples[names[1]], osamples[names[2]]]
    key = f"{names[1]}.vs.{names[2]}"
    rot = ref["pairs"][key]["CPC"][:, :20]
    op = oracle.project_pair(pr, ref["pairs"][key]["genes"], rot, var_scale, common_centering)
    for side in (0, 1):
        nr, nc = C.c_int32(0), C.c_int32(0)
        ctx.check(ctx.lib.cg_pair_projection(ctx.h, 2, side, C.byref(nr), C.byref(nc), None))
        p = np.zeros((nr.value, nc.value), order="F")
        ctx.check(ctx.lib.cg_pair_projection(ctx.h, 2, side, None, None, p.ctypes.data_as(C.POINTER(C.c_double))))
        np.testing.assert_allclose(p, op[side], rtol=1e-9, atol=1e-11)
    _graphs_equal(g, ref["graph"], rtol=1e-6)
    adj = oracle.adjacency_csr(ref["graph"])
    a = g.as_adj()
    assert np.array_equal(a.indptr, adj.indptr) and np.array_equal(a.indices, adj.indices)
    np.testing.assert_allclose(a.data, adj.data, rtol=1e-6)
    assert g.names[0] == small_panel[names[0]].cells[g.vertices[0]]


def test_local_edges_and_merge(ctx, oracle, small_panel):
    """k.self edges: self + exact-zero distances dropped, reciprocal edges summed by simplify."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    names = list(small_panel)
    frm, to, w, typ = [], [], [], []
    off = 0
    for n in names:
        nb, q, ww = oracle.local_edges(osamples[n].pca, 5, 0.1)
        frm.append(nb + off); to.append(q + off); w.append(ww); typ.append(np.zeros(len(ww), np.int32))
        off += osamples[n].counts.shape[0]
    og = oracle.assemble_graph(np.concatenate(frm).astype(np.int64), np.concatenate(to).astype(np.int64),
                               np.concatenate(w), np.concatenate(typ))
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con._ensure_panel(names)
    P = conos_b200._lib.CgParams()
    ctx.lib.cg_params_default(C.byref(P))
    P.k_self = 5
    edges = C.c_void_p()
    sidx = np.arange(len(names), dtype=np.int32)
    ctx.check(ctx.lib.cg_local_edges(ctx.h, con._panel, sidx.ctypes.data_as(C.POINTER(C.c_int32)), len(names),
                                     C.byref(P), C.byref(edges), None))
    g = con._assemble(edges, names)
    ctx.lib.cg_edges_free(edges)
    _graphs_equal(g, og)
    assert (g.type == 0).all()


def test_local_edges_with_100_pcs_and_wide_common_space(ctx, oracle):
    """Per-sample PCA with pagoda2's default 100 columns (local neighbours through the exact SIMT kNN kernel) and a
    common space of 70 components (PCA, r = 35 per sample: FP64 eigen driver, projection and kNN with 128-wide
    panels): the whole graph against the oracle with the oracle's rotations."""
    import conos_b200
    from conos_b200.panel import make_panel
    panel = make_panel(13, 2, [500, 420], n_genes=300, n_types=5, n_pcs=100, use_torch=False)
    osamples = to_oracle_samples(oracle, panel)
    ref = oracle.build_graph(osamples, k=10, k_self=6, space="PCA", ncomps=70, n_odgenes=300)
    con = conos_b200.Conos(panel, ctx=ctx)
    con.pairs["PCA"] = {key: {"CPC": red["CPC"], "genes": red["genes"]} for key, red in ref["pairs"].items()}
    g = con.buildGraph(k=10, k_self=6, space="PCA", ncomps=70, n_odgenes=300)
    _graphs_equal(g, ref["graph"], rtol=1e-6)
    # and with the rotation computed on the device (exact mode): 35 eigenvectors per sample
    con2 = conos_b200.Conos(panel, ctx=ctx)
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0)
    try:
        g2 = con2.buildGraph(k=10, k_self=6, space="PCA", ncomps=70, n_odgenes=300)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    key0 = list(ref["pairs"])[0]
    assert con2.pairs["PCA"][key0]["CPC"].shape == ref["pairs"][key0]["CPC"].shape
    assert con2.pairs["PCA"][key0]["CPC"].shape[1] == 70
    a = {(x, y) for x, y in zip(g2.vertices[g2.edge_from].tolist(), g2.vertices[g2.edge_to].tolist())}
    b = {(x, y) for x, y in zip(g.vertices[g.edge_from].tolist(), g.vertices[g.edge_to].tolist())}
    assert len(a & b) / len(a | b) > 0.98
    con.close()
    con2.close()


def test_assemble_graph_semantics(ctx, oracle):
    """w <= 0 rows dropped before vertex numbering, loops removed, multi-edges summed in table order, type = first."""
    rng = np.random.default_rng(4)
    n_cells = 200
    frm = rng.integers(0, n_cells, 5000).astype(np.int32)
    to = rng.integers(0, n_cells, 5000).astype(np.int32)
    w = rng.uniform(-0.2, 1.0, 5000)
    typ = rng.integers(0, 2, 5000).astype(np.int32)
    og = oracle.assemble_graph(frm.astype(np.int64), to.astype(np.int64), w, typ)
    lib, h = ctx.lib, ctx.h
    e = C.c_void_p()
    ip = C.POINTER(C.c_int32)
    ctx.check(lib.cg_edges_append_host(h, C.byref(e), frm.ctypes.data_as(ip), to.ctypes.data_as(ip),
                                       w.ctypes.data_as(C.POINTER(C.c_double)), typ.ctypes.data_as(ip), len(w)))
    gh = C.c_void_p()
    ctx.check(lib.cg_assemble_graph(h, e, n_cells, C.byref(gh)))
    nv, ne = C.c_int32(0), C.c_int64(0)
    lib.cg_graph_sizes(gh, C.byref(nv), C.byref(ne))
    assert nv.value == len(og["vertices"]) and ne.value == len(og["from"])
    v = np.empty(nv.value, np.int32); f = np.empty(ne.value, np.int32); t = np.empty(ne.value, np.int32)
    ww = np.empty(ne.value); ty = np.empty(ne.value, np.int32)
    ctx.check(lib.cg_graph_fetch(h, gh, v.ctypes.data_as(ip), f.ctypes.data_as(ip), t.ctypes.data_as(ip),
                                 ww.ctypes.data_as(C.POINTER(C.c_double)), ty.ctypes.data_as(ip)))
    assert np.array_equal(v, og["vertices"]) and np.array_equal(f, og["from"]) and np.array_equal(t, og["to"])
    assert np.array_equal(ty, og["type"])
    np.testing.assert_allclose(ww, og["weight"], rtol=1e-14)
    lib.cg_graph_free(gh)
    lib.cg_edges_free(e)


def test_empty_and_error_paths(ctx):
    from conos_b200 import ops, ConosGpuError
    with pytest.raises(ValueError):
        ops.crossKnn(np.zeros((3, 2)), np.zeros((3, 3)), 2, ctx=ctx)
    with pytest.raises(ValueError, match="distance metrics"):
        ops.crossKnn(np.zeros((3, 2)), np.zeros((3, 2)), 2, metric="cosine", ctx=ctx)
    with pytest.raises(ConosGpuError, match="k1 must be >= k"):
        ops.getNeighborMatrix(np.ones((4, 2)), np.ones((4, 2)), 3, k1=2, ctx=ctx)


def test_rebalancing_kernels(ctx, oracle):
    from conos_b200 import ops
    rng = np.random.default_rng(8)
    n_cells, n_levels, ne = 300, 4, 4000
    ri = rng.integers(0, n_cells, ne).astype(np.int32)
    ci = rng.integers(0, n_cells, ne).astype(np.int32)
    w = rng.uniform(0.01, 1, ne)
    fl = rng.integers(1, n_levels + 1, n_cells).astype(np.int32)
    fl[:n_levels] = np.arange(1, n_levels + 1)
    for normalize in (True, False):
        m = ops.getSumWeightMatrix(w, ri, ci, fl, normalize, ctx=ctx)
        np.testing.assert_allclose(m, oracle.sum_weight_matrix(w, ri, ci, fl, normalize), rtol=1e-12, atol=1e-15)
    div = ops.getSumWeightMatrix(w, ri, ci, fl, True, ctx=ctx) + 0.05
    np.testing.assert_allclose(ops.adjustWeightsByCellBalancingC(w, ri, ci, fl, div, ctx=ctx),
                               oracle.adjust_weights_c(w, ri, ci, fl, div), rtol=1e-14)


def test_balance_edge_weights(ctx, oracle, small_panel):
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300,
                             balance_edge_weights=True)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.pairs["PCA"] = {key: {"CPC": red["CPC"], "genes": red["genes"]} for key, red in ref["pairs"].items()}
    g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, balance_edge_weights=True)
    ri, ci, w = ref["balanced"]
    a = sp.csc_matrix((w, (ri, ci)), shape=(len(g.vertices),) * 2).tocsr()
    a.sort_indices()
    b = g.as_adj()
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    np.testing.assert_allclose(b.data, a.data, rtol=1e-5)


@pytest.mark.parametrize("mode", ["per_sample", "per_cell"])
def test_balancing_factors_and_k_same_factor(ctx, oracle, small_panel, mode):
    """balancing.factor.per.sample (with k.same.factor: k.cur of the pairs inside one level, R/conclass.R:230-233),
    balancing.factor.per.cell and same.factor.downweight (R/conclass.R:346-368, R/conos.R:936-967) through
    buildGraph: the 50 balancing rounds run as one device call."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    names = list(small_panel)
    kw = dict(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, same_factor_downweight=0.5)
    if mode == "per_sample":
        extra = dict(balancing_factor_per_sample={names[0]: "x", names[1]: "x", names[2]: "y"}, k_same_factor=4,
                     balance_edge_weights=True)
    else:
        per_cell = {c: ("a" if t % 3 else "b") for n in names for t, c in enumerate(small_panel[n].cells)}
        extra = dict(balancing_factor_per_cell=per_cell, balance_edge_weights=False)
    ref = oracle.build_graph(osamples, **kw, **extra)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.pairs["PCA"] = {key: {"CPC": red["CPC"], "genes": red["genes"]} for key, red in ref["pairs"].items()}
    g = con.buildGraph(**kw, **extra)
    ri, ci, w = ref["balanced"]
    a = sp.csc_matrix((w, (ri, ci)), shape=(len(g.vertices),) * 2).tocsr()
    a.sort_indices()
    b = g.as_adj()
    assert np.array_equal(g.vertices, ref["graph"]["vertices"])
    assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices)
    np.testing.assert_allclose(b.data, a.data, rtol=1e-6)
    # the edge list follows the adjacency matrix, the type attribute is gone (graph rebuilt from the matrix)
    up = sp.triu(b, k=1).tocoo()
    o = np.lexsort((up.col, up.row))
    assert np.array_equal(g.edge_from, up.row[o]) and np.array_equal(g.edge_to, up.col[o])
    np.testing.assert_allclose(g.weight, up.data[o], rtol=0, atol=0)
    assert (g.type == -1).all()
    con.close()


def test_reference_wij(ctx, oracle, small_panel):
    import conos_b200
    from conos_b200 import ops
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
    adj = oracle.adjacency_csr(ref["graph"])
    want = oracle.build_wij_matrix(adj, perplexity=10)
    got = ops.buildWijMatrix(adj, perplexity=10, ctx=ctx)
    want.sort_indices(); got.sort_indices()
    assert np.array_equal(want.indptr, got.indptr) and np.array_equal(want.indices, got.indices)
    np.testing.assert_allclose(got.data, want.data, rtol=1e-9)


def test_spcov_and_cpca(ctx, oracle, small_panel):
    from conos_b200 import ops
    osamples = list(to_oracle_samples(oracle, small_panel).values())[:2]
    od = oracle.common_overdispersed_genes(osamples, 300)
    sm = oracle.scaled_matrices(osamples, od, True)
    covs = []
    for x in sm:
        cm = np.asarray(x.mean(axis=0)).ravel()
        want = oracle.spcov(x, cm)
        got = ops.spcov(x, cm, ctx=ctx)
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)
        covs.append(want)
    ncells = [x.shape[0] for x in sm]
    want, ev = oracle.cpca_fast(covs, ncells, ncomp=8, maxit=500, tol=1e-5)
    got = ops.cpcaF(np.stack(covs, axis=2), ncells, ncomp=8, maxit=500, tol=1e-5, eigenvR=ev, ctx=ctx)
    assert np.array_equal(got["niter"], want["niter"])
    np.testing.assert_allclose(got["CPC"], want["CPC"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(got["D"], want["D"], rtol=1e-9)


def _principal_cosines(a, b):
    qa, _ = np.linalg.qr(a)
    qb, _ = np.linalg.qr(b)
    return np.linalg.svd(qa.T @ qb, compute_uv=False)


@pytest.mark.parametrize("exact", [True, False])
@pytest.mark.parametrize("ncomps,score", [(20, True), (30, False)])
def test_device_pca_rotation(ctx, oracle, small_panel, ncomps, score, exact):
    """quickPlainPCA on the device: per-sample subspaces equal the exact ones (irlba itself is only 1e-5 accurate),
    columns interleaved A1,B1,A2,..., nv = variance fractions; the downstream graph matches the oracle's."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=ncomps, n_odgenes=300,
                             score_component_variance=score)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    # exact: FP64 CUDA-core reduction (tight parity); default: bf16x3 tensor-core Gram + Chebyshev iteration to
    # irlba's own tolerance (1e-5 relative residual), so only irlba-level agreement can be asked for
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0 if exact else 0.0)
    try:
        g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=ncomps, n_odgenes=300, score_component_variance=score)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    cos_tol, dot_tol, nv_tol, jac_tol = (1e-10, 1e-8, 1e-8, 0.999) if exact else (1e-6, None, 1e-4, 0.995)
    r = ncomps // 2
    for key, red in ref["pairs"].items():
        mine = con.pairs["PCA"][key]
        assert mine["genes"] == red["genes"]
        assert mine["CPC"].shape == red["CPC"].shape
        for side in (0, 1):
            cs = _principal_cosines(mine["CPC"][:, side::2], red["CPC"][:, side::2])
            assert cs.min() > 1 - cos_tol, (key, side, cs.min())
            if dot_tol is not None:
                # individual eigenvectors (up to sign): well separated spectrum in this panel
                dots = np.abs(np.sum(mine["CPC"][:, side::2] * red["CPC"][:, side::2], axis=0))
                assert dots.min() > 1 - dot_tol
        if score:
            for side in (0, 1):
                np.testing.assert_allclose(mine["nv"][side], red["nv"][side], rtol=nv_tol)
    og = ref["graph"]
    # the rotations agree to ~1e-9, so at most a handful of near-tied neighbours may differ
    mine_e = set(zip(g.vertices[g.edge_from].tolist(), g.vertices[g.edge_to].tolist()))
    ref_e = set(zip(og["vertices"][og["from"]].tolist(), og["vertices"][og["to"]].tolist()))
    mine_e = {(min(a, b), max(a, b)) for a, b in mine_e}
    ref_e = {(min(a, b), max(a, b)) for a, b in ref_e}
    jacc = len(mine_e & ref_e) / len(mine_e | ref_e)
    print(f"JACCARD pca exact={exact} ncomps={ncomps}: {jacc:.6f}")
    assert jacc > jac_tol, jacc


@pytest.mark.parametrize("matching,k,k1", [("NN", 10, 10), ("mNN", 40, 40), ("mNN", 8, 60), ("NN", 6, 25), ("mNN", 15, 200)])
def test_neighbor_matrix_general_paths(ctx, oracle, matching, k, k1):
    """matching='NN', k > 32, and k1 > k (reduceEdgesInGraphIteratively + pareDownHubEdges) against the oracle."""
    from conos_b200 import ops
    rng = np.random.default_rng(k * 100 + k1)
    p1, p2 = mixture(rng, 500, 20, k=5), mixture(rng, 380, 20, k=5)
    i, j, x = ops.getNeighborMatrix(p1, p2, k, k1=k1, matching=matching, ctx=ctx)
    oi, oj, ox = oracle.neighbor_matrix(p1, p2, k, k1=k1, matching=matching)
    assert np.array_equal(i, oi) and np.array_equal(j, oj)
    np.testing.assert_allclose(x, ox, rtol=1e-12)
    if k1 > k:
        # the greedy reduction is a heuristic: degrees end near k, well below the k1 it started from
        assert max(np.bincount(i).max(), np.bincount(j).max()) <= 2 * k + 5


def test_pare_down_hub_edges_entry_point(ctx, oracle):
    from conos_b200 import ops
    m = sp.random(300, 220, density=0.15, random_state=7, format="csc")
    m.sort_indices()
    for k, klow in [(5, -1), (8, 3), (40, 2)]:
        rn_o = np.bincount(m.indices, minlength=300).astype(np.int32)
        rn_g = rn_o.copy()
        want = oracle.pare_down_hub_edges(m, rn_o, k, klow)
        got = ops.pareDownHubEdges(m, rn_g, k, klow, ctx=ctx)
        assert np.array_equal(got, want) and np.array_equal(rn_g, rn_o)


def test_build_graph_alignment_strength(ctx, oracle, small_panel):
    """alignment.strength > 0: k1 = max(k, round(max_cells * a^2)), cor.base = 1 + min(1, 10a), hub paring."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, alignment_strength=0.3)
    assert ref["k1"] == 38 and ref["cor_base"] == 2.0
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.pairs["PCA"] = {key: {"CPC": red["CPC"], "genes": red["genes"]} for key, red in ref["pairs"].items()}
    g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300, alignment_strength=0.3)
    _graphs_equal(g, ref["graph"], rtol=1e-6)


@pytest.mark.parametrize("exact", [True, False])
def test_device_cpca_rotation(ctx, oracle, small_panel, exact):
    """quickCPCA on the device: covariances, irlba-replacing initial vectors and the cpcaF iteration; the common
    components match the oracle's (exact mode tightly; the tensor-core Gram only to its 1e-6 accuracy)."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="CPCA", ncomps=12, n_odgenes=300,
                             score_component_variance=True)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0 if exact else 0.0)
    try:
        g = con.buildGraph(k=10, k_self=5, space="CPCA", ncomps=12, n_odgenes=300, score_component_variance=True)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    tol = 1e-6 if exact else 2e-3
    for key, red in ref["pairs"].items():
        mine = con.pairs["CPCA"][key]
        assert mine["CPC"].shape == red["CPC"].shape and mine["CPC"].shape[1] == 12
        dots = np.abs(np.sum(mine["CPC"] * red["CPC"], axis=0))
        assert dots.min() > 1 - tol, (key, dots.min())
        for side in (0, 1):
            np.testing.assert_allclose(mine["nv"][side], red["nv"][side], rtol=100 * tol)
        # the other fields of cpcaF's result (src/cpca.cpp:93-100): D (k x ncomp) and niter
        assert mine["D"].shape == red["D"].shape == (12, 2) and mine["niter"].shape == (12,)
        np.testing.assert_allclose(mine["D"], red["D"], rtol=1e-6 if exact else 1e-2)
        if exact:
            assert np.array_equal(mine["niter"], red["niter"]), (key, mine["niter"], red["niter"])
        else:
            assert np.abs(mine["niter"] - red["niter"]).max() <= 3
    og = ref["graph"]
    mine_e = {(min(a, b), max(a, b)) for a, b in zip(g.vertices[g.edge_from].tolist(), g.vertices[g.edge_to].tolist())}
    ref_e = {(min(a, b), max(a, b))
             for a, b in zip(og["vertices"][og["from"]].tolist(), og["vertices"][og["to"]].tolist())}
    jacc = len(mine_e & ref_e) / len(mine_e | ref_e)
    print(f"JACCARD cpca exact={exact}: {jacc:.6f}")
    assert jacc > 0.995   # measured 1.000000 in both modes (round 1's default mode: > 0.97)


def test_cca_with_oracle_variates(ctx, oracle, small_panel):
    """space='CCA' with the oracle's u, v supplied (cache-hit path): everything after the reduction must agree,
    including cells dropped for empty rows and alignment.strength > 0 (k1 > k, cor.base = 2)."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="CCA", ncomps=10, n_odgenes=300, alignment_strength=0.2)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.pairs["CCA"] = {key: {"u": red["u"], "v": red["v"], "rows_u": np.nonzero(red["keep"][0])[0],
                              "rows_v": np.nonzero(red["keep"][1])[0]} for key, red in ref["pairs"].items()}
    g = con.buildGraph(k=10, k_self=5, space="CCA", ncomps=10, n_odgenes=300, alignment_strength=0.2)
    _graphs_equal(g, ref["graph"], rtol=1e-6)


@pytest.mark.parametrize("exact", [True, False])
def test_device_cca_reduction(ctx, oracle, small_panel, exact):
    """quickCCA on the device (implicit n1 x n1 operator instead of the dense n1 x n2 cross product).  With the
    FP64 Gram and a 1e-10 eigen tolerance the variates match the dense SVD tightly; the default mode (tensor-core
    Gram, irlba's own 1e-5 tolerance) is held to irlba-level agreement."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="CCA", ncomps=10, n_odgenes=300)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0 if exact else 0.0)
    try:
        g = con.buildGraph(k=10, k_self=5, space="CCA", ncomps=10, n_odgenes=300)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    rt, at, jt = (1e-8, 1e-7, 0.995) if exact else (1e-3, 5e-3, 0.995)   # measured 1.000000 in both modes
    for key, red in ref["pairs"].items():
        mine = con.pairs["CCA"][key]
        assert np.array_equal(mine["rows_u"], np.nonzero(red["keep"][0])[0])
        assert np.array_equal(mine["rows_v"], np.nonzero(red["keep"][1])[0])
        np.testing.assert_allclose(mine["d"], red["d"], rtol=rt)
        for a, b in ((mine["u"], red["u"]), (mine["v"], red["v"])):
            if exact:
                sign = np.sign(np.sum(a * b, axis=0))
                np.testing.assert_allclose(a * sign[None, :], b, rtol=0, atol=at * np.abs(b).max())
            else:
                # irlba's tolerance (residual <= 1e-5 sigma_1^2) fixes the canonical subspace, not the individual
                # variates of nearly equal correlations
                assert _principal_cosines(a, b).min() > 1 - 1e-3
        # u_j and v_j flip together
        assert np.array_equal(np.sign(np.sum(mine["u"] * red["u"], axis=0)), np.sign(np.sum(mine["v"] * red["v"], axis=0)))
        # gene loadings and variance fractions (R/conos.R:349-357)
        assert mine["ul"].shape == red["ul"].shape and mine["vl"].shape == red["vl"].shape
        np.testing.assert_allclose(np.linalg.norm(mine["ul"], axis=0), 1.0, rtol=1e-12)
        if exact:
            for a, b in ((mine["ul"], red["ul"]), (mine["vl"], red["vl"])):
                sign = np.sign(np.sum(a * b, axis=0))
                np.testing.assert_allclose(a * sign[None, :], b, rtol=0, atol=1e-6)
            for side in (0, 1):
                np.testing.assert_allclose(mine["nv"][side], red["nv"][side], rtol=1e-6)
        else:
            assert _principal_cosines(mine["ul"], red["ul"]).min() > 1 - 1e-3
            for side in (0, 1):
                np.testing.assert_allclose(np.sort(mine["nv"][side]), np.sort(red["nv"][side]), rtol=5e-2)
    og = ref["graph"]
    mine_e = {(min(a, b), max(a, b)) for a, b in zip(g.vertices[g.edge_from].tolist(), g.vertices[g.edge_to].tolist())}
    ref_e = {(min(a, b), max(a, b))
             for a, b in zip(og["vertices"][og["from"]].tolist(), og["vertices"][og["to"]].tolist())}
    jacc = len(mine_e & ref_e) / len(mine_e | ref_e)
    print(f"JACCARD cca exact={exact}: {jacc:.6f}")
    assert jacc > jt


@pytest.mark.parametrize("n_types,ncomps", [(2, 30), (6, 20)])
def test_device_pca_wide_spectrum_default_mode(ctx, n_types, ncomps):
    """Fewer cell types than components: the wanted eigenvalues span more than a decade and the tail sits in the
    noise bulk.  The tensor-core solver (bf16x3 GEMMs, fp32 panels) must ramp its Chebyshev degree instead of losing
    the small directions; its subspaces agree with the FP64 path to irlba's own accuracy.  Samples express
    different gene subsets, so the pair gene sets are proper, reordered subsets of each sample's genes."""
    import conos_b200
    from conos_b200.panel import make_panel
    panel = make_panel(11, 3, [900, 800, 700], n_genes=400, n_types=n_types, n_pcs=20, use_torch=False)
    res = {}
    try:
        for exact in (1, 0):
            ctx.lib.cg_set_option(ctx.h, b"exact_reduction", float(exact))
            con = conos_b200.Conos(panel, ctx=ctx)
            con.buildGraph(k=10, k_self=5, space="PCA", ncomps=ncomps, n_odgenes=400, score_component_variance=True)
            res[exact] = con.pairs["PCA"]
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    for key in res[1]:
        a, b = res[1][key], res[0][key]
        assert a["genes"] == b["genes"]
        for side in (0, 1):
            cs = _principal_cosines(a["CPC"][:, side::2], b["CPC"][:, side::2])
            assert cs.min() > 1 - 1e-5, (key, side, cs.min())
            np.testing.assert_allclose(b["nv"][side], a["nv"][side], rtol=1e-3)


def test_device_pca_large_gene_universe_default_mode(ctx):
    """A gene universe above 4096 (the union of many samples' od lists on real data): the tensor-core Gram matrix and
    eigen solver still serve it (no FP64 fallback), and the subspaces agree with the FP64 path."""
    import ctypes as C
    import conos_b200
    from conos_b200.panel import make_panel
    panel = make_panel(17, 2, [1500, 1300], n_genes=6000, n_types=5, n_pcs=20, use_torch=False)
    res = {}
    try:
        for exact in (1, 0):
            ctx.lib.cg_set_option(ctx.h, b"exact_reduction", float(exact))
            con = conos_b200.Conos(panel, ctx=ctx)
            ctx.lib.cg_profile(ctx.h, 1)
            con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=6000)
            buf = C.create_string_buffer(8192)
            ctx.lib.cg_profile_dump(ctx.h, buf, 8192)
            ctx.lib.cg_profile(ctx.h, 0)
            prof = dict(kv.split("=") for kv in buf.value.decode().split(";") if kv)
            res[exact] = (con.pairs["PCA"], prof)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
        ctx.lib.cg_profile(ctx.h, 0)
    # the default mode went through the tensor-core solver: two problems (one per side), none handed to FP64
    assert float(res[0][1].get("reduction.eig.problems", 0)) == 2.0, res[0][1]
    assert float(res[0][1].get("reduction.eig.fp64_redo_pairs", 0)) == 0.0
    for key in res[1][0]:
        a, b = res[1][0][key], res[0][0][key]
        assert a["CPC"].shape[0] > 4096
        for side in (0, 1):
            cs = _principal_cosines(a["CPC"][:, side::2], b["CPC"][:, side::2])
            assert cs.min() > 1 - 1e-5, (key, side, cs.min())


@pytest.mark.parametrize("exact", [True, False])
def test_reference_fixture_on_gpu(ctx, golden, exact):
    """buildGraph(ncomps=25) on the reference's own fixture, as tests/testthat/test_functions.R:24 calls it, against
    the committed oracle vectors: one vertex per cell (59), and -- with the FP64 reductions -- the oracle's edge
    set with its weights; the tensor-core reductions agree to irlba's accuracy, so a few near-tied neighbours
    may differ."""
    import conos_b200
    mod, vec = golden
    panel = mod.load_reference_panel(conos_b200.Pagoda2)
    con = conos_b200.Conos(panel, ctx=ctx)
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0 if exact else 0.0)
    try:
        g = con.buildGraph(ncomps=25)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    assert g.vcount() == 59
    mine = {(min(a, b), max(a, b)): w for a, b, w in zip(g.vertices[g.edge_from].tolist(),
                                                        g.vertices[g.edge_to].tolist(), g.weight.tolist())}
    gv = vec["graph_vertices"]
    ref = {(min(a, b), max(a, b)): w for a, b, w in zip(gv[vec["graph_from"]].tolist(), gv[vec["graph_to"]].tolist(),
                                                       vec["graph_weight"].tolist())}
    common = set(mine) & set(ref)
    jac = len(common) / len(set(mine) | set(ref))
    if exact:
        assert np.array_equal(g.vertices, gv)
        assert set(mine) == set(ref)
        np.testing.assert_allclose([mine[e] for e in sorted(common)], [ref[e] for e in sorted(common)], rtol=1e-6)
    else:
        print(f"JACCARD reference fixture exact={exact}: {jac:.6f}")
        assert jac > 0.995, jac   # measured 1.000000 (round 1: > 0.9)


def test_unconverged_tensor_core_pairs_go_to_fp64(ctx, small_panel):
    """A pair whose tensor-core eigen solve has not met the tolerance after the allowed sweeps is recomputed by the
    FP64 driver instead of being returned half-converged (forced here by allowing a single sweep)."""
    import conos_b200
    res = {}
    try:
        for sweeps in (40, 1):
            ctx.check(ctx.lib.cg_set_option(ctx.h, b"tc_eig_max_sweeps", float(sweeps)))
            con = conos_b200.Conos(small_panel, ctx=ctx)
            con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
            res[sweeps] = con.pairs["PCA"]
    finally:
        ctx.lib.cg_set_option(ctx.h, b"tc_eig_max_sweeps", 40.0)
    for key in res[40]:
        for side in (0, 1):
            cs = _principal_cosines(res[40][key]["CPC"][:, side::2], res[1][key]["CPC"][:, side::2])
            assert cs.min() > 1 - 1e-5, (key, side, cs.min())


def test_baseline_config0_on_gpu(ctx, golden):
    """BASELINE.json configs[0]: the bundled panel, buildGraph(k=15, k.self=5, space='PCA', ncomps=30), FP64 reductions:
    the oracle's committed graph, vertex order, edges and weights."""
    import conos_b200
    mod, vec = golden
    con = conos_b200.Conos(mod.load_reference_panel(conos_b200.Pagoda2), ctx=ctx)
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0)
    try:
        g = con.buildGraph(k=15, k_self=5, space="PCA", ncomps=30)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    assert np.array_equal(g.vertices, vec["c0_vertices"])
    assert np.array_equal(g.edge_from, vec["c0_from"]) and np.array_equal(g.edge_to, vec["c0_to"])
    assert np.array_equal(g.type, vec["c0_type"])
    np.testing.assert_allclose(g.weight, vec["c0_weight"], rtol=1e-6)


def test_gene_universe_subsetting(ctx, oracle, small_panel):
    """n.odgenes below the number of genes: only the union of the samples' top-n od lists is uploaded (real pagoda2
    objects carry tens of thousands of genes); gene columns are remapped and the graph equals the oracle's."""
    import conos_b200
    osamples = to_oracle_samples(oracle, small_panel)
    ref = oracle.build_graph(osamples, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=120)
    con = conos_b200.Conos(small_panel, ctx=ctx)
    ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 1.0)
    try:
        g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=120)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"exact_reduction", 0.0)
    assert any(pc is not None for pc in con._panel_col.values()), "the panel was not subset: the test does not bite"
    for key, red in ref["pairs"].items():
        assert con.pairs["PCA"][key]["genes"] == red["genes"]
    _graphs_equal(g, ref["graph"], rtol=1e-6)


def test_tensor_core_projection_matches_fp64_projection(ctx, small_panel):
    """Default mode projects device-computed rotations with one bf16x3 tensor-core GEMM per sample (fp32 accumulation);
    with "tc_projection" = 0 the same rotations go through the FP64 sparse product.  The rotations are identical in
    both runs, so the graphs differ only where a neighbour was tied to ~1e-6: the same edges up to a handful, and
    the same weights on the common ones (getPcaBasedNeighborMatrix, R/conos.R:776-833)."""
    import conos_b200
    graphs = {}
    try:
        for tc in (1, 0):
            ctx.check(ctx.lib.cg_set_option(ctx.h, b"tc_projection", float(tc)))
            con = conos_b200.Conos(small_panel, ctx=ctx)
            g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
            graphs[tc] = {(a, b): w for a, b, w in zip(g.edge_from.tolist(), g.edge_to.tolist(), g.weight.tolist())}
            rot = con.pairs["PCA"]
            graphs[(tc, "rot")] = {key: red["CPC"] for key, red in rot.items()}
    finally:
        ctx.lib.cg_set_option(ctx.h, b"tc_projection", 1.0)
    for key, r in graphs[(1, "rot")].items():
        assert np.array_equal(r, graphs[(0, "rot")][key])
    a, b = graphs[1], graphs[0]
    common = set(a) & set(b)
    assert len(common) / len(set(a) | set(b)) > 0.995
    np.testing.assert_allclose([a[e] for e in sorted(common)], [b[e] for e in sorted(common)], rtol=2e-4, atol=1e-6)


def test_options_and_counters(ctx, small_panel):
    """cg_set_option rejects unknown names and out-of-range values; cg_get_stat exposes the counters and the two
    memory pools (include/conos_gpu.h); the resident panel lives in its own pool."""
    import conos_b200
    assert ctx.lib.cg_set_option(ctx.h, b"no_such_option", 1.0) != 0
    assert ctx.lib.cg_set_option(ctx.h, b"tc_eig_degree", 1.0) != 0
    assert ctx.lib.cg_set_option(ctx.h, b"tc_eig_degree", 10.0) == 0
    with pytest.raises(KeyError):
        ctx.stat("no_such_counter")
    launches0 = ctx.stat("kernel_launches")
    con = conos_b200.Conos(small_panel, ctx=ctx)
    con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
    assert ctx.stat("kernel_launches") > launches0
    assert ctx.stat("panel_pool_reserved_bytes") > 0 and ctx.stat("scratch_pool_reserved_bytes") > 0
    assert ctx.stat("knn_redo_blocks") >= 0 and ctx.stat("eig_redo_pairs") >= 0
