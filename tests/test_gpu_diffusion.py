"""GPU parity of the diffusion over the joint graph (smooth_count_matrix / propagate_labels,
src/propagate_labels.cpp:65-222) against the oracle port, through the C ABI.  Every row is summed in the reference's
edge order with separate multiply and add, so the only difference is exp() (libm vs device, 1 ulp): values to 1e-12
relative, iteration counts equal."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def random_multigraph(rng, n, n_edges):
    """Edges in arbitrary order with duplicates, both orientations and self loops, as an edge list may hold them."""
    ef = rng.integers(0, n, n_edges).astype(np.int32)
    et = rng.integers(0, n, n_edges).astype(np.int32)
    et[:5] = ef[:5]                       # self loops
    ef[5:10], et[5:10] = et[10:15], ef[10:15]   # the same pair in the other orientation
    w = rng.uniform(0.05, 1.0, n_edges)
    return ef, et, w


@pytest.mark.parametrize("n,n_edges,n_cols,normalize,fixed", [(500, 4000, 7, True, True), (500, 4000, 70, False, False),
                                                              (1200, 30000, 33, True, False), (40, 60, 1, False, True)])
def test_smooth_count_matrix_matches_oracle(ctx, oracle, n, n_edges, n_cols, normalize, fixed):
    from conos_b200 import ops
    rng = np.random.default_rng(n + n_cols)
    ef, et, w = random_multigraph(rng, n, n_edges)
    m = rng.gamma(0.3, 2.0, (n, n_cols))
    fx = (rng.random(n) < 0.2) if fixed else None
    for iters, tol in ((1, 1e-3), (6, 1e-3), (30, 0.05)):
        om, oit, onrm = oracle.smooth_count_matrix(ef, et, w, m, fx, max_n_iters=iters, diffusion_fading=2.0,
                                                   diffusion_fading_const=0.1, tol=tol, normalize=normalize,
                                                   return_info=True)
        gm, git, gnrm = ops.smooth_count_matrix(ef, et, w, m, fx, max_n_iters=iters, diffusion_fading=2.0,
                                                diffusion_fading_const=0.1, tol=tol, normalize=normalize, ctx=ctx,
                                                return_info=True)
        assert git == oit, (iters, tol, git, oit)
        np.testing.assert_allclose(gm, om, rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(gnrm, onrm, rtol=1e-9)
        if fx is not None:
            assert np.array_equal(gm[fx], m[fx])          # fixed rows never move


def test_smooth_count_matrix_stops_before_the_converged_iterate(ctx, oracle):
    """The reference tests the norm BEFORE taking the new matrix (src/propagate_labels.cpp:166-170): with a tolerance
    the third iterate meets, the second one is returned and the reported iteration index is 2."""
    from conos_b200 import ops
    rng = np.random.default_rng(3)
    ef, et, w = random_multigraph(rng, 300, 3000)
    m = rng.random((300, 5))
    norms = [oracle.smooth_count_matrix(ef, et, w, m, max_n_iters=t, tol=0.0, normalize=True, return_info=True)[2]
             for t in (1, 2, 3)]
    tol = 0.5 * (norms[1] + norms[2])
    assert norms[2] < tol < norms[1]
    two = oracle.smooth_count_matrix(ef, et, w, m, max_n_iters=2, tol=0.0, normalize=True)
    gm, git, _ = ops.smooth_count_matrix(ef, et, w, m, max_n_iters=50, tol=tol, normalize=True, ctx=ctx, return_info=True)
    assert git == 2
    np.testing.assert_allclose(gm, two, rtol=1e-12)


def test_propagate_labels_on_the_joint_graph(ctx, oracle, small_panel):
    """Conos$propagateLabels(method='diffusion') on a graph built by buildGraph: label distribution, labels and
    uncertainty against the oracle on the same edge list."""
    import conos_b200
    con = conos_b200.Conos(small_panel, ctx=ctx)
    g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
    names = np.asarray(g.names, dtype=object)
    rng = np.random.default_rng(5)
    known = rng.choice(len(names), len(names) // 5, replace=False)
    # labels in the caller's order; label ids follow first appearance (order_strings, :33-38)
    lab_of = {names[t]: f"type{(t * 7) % 4}" for t in known}
    lab_of["not_a_cell_of_the_graph"] = "ghost"
    for fixed in (True, False):
        res = con.propagateLabels(lab_of, max_iters=40, tol=1e-3, fixed_initial_labels=fixed)
        ids = {}
        lab = np.full(len(names), -1, np.int32)
        pos = {n: t for t, n in enumerate(names)}
        for cell, l in lab_of.items():
            if cell in pos:
                lab[pos[cell]] = ids.setdefault(l, len(ids))
        assert res["label_names"] == list(ids) and "ghost" not in ids
        od, oit, _ = oracle.propagate_labels(g.edge_from, g.edge_to, g.weight, len(names), lab, len(ids), max_n_iters=40,
                                             diffusion_fading=10.0, diffusion_fading_const=0.1, tol=1e-3,
                                             fixed_initial_labels=fixed, return_info=True)
        np.testing.assert_allclose(res["label_distribution"], od, rtol=1e-11, atol=1e-300, equal_nan=True)
        ok = ~np.isnan(od).any(axis=1)
        assert ok.sum() > 0.9 * len(names)
        assert [res["label_names"].index(l) for l in res["labels"][ok]] == list(np.argmax(od[ok], axis=1))
        np.testing.assert_allclose(res["uncertainty"][ok], 1 - od[ok].max(axis=1), rtol=1e-11, atol=1e-14)
        if fixed:
            assert all(res["labels"][pos[c]] == l for c, l in lab_of.items() if c in pos)
    con.close()


def test_graph_handle_and_host_edge_list_agree(ctx):
    """cg_smooth_matrix / cg_propagate_labels on a device graph handle (its simplified edges never leave the device) give
    bit for bit what the fetched edge list gives."""
    from conos_b200 import ops
    rng = np.random.default_rng(8)
    n_cells = 400
    frm = rng.integers(0, n_cells, 6000).astype(np.int32)
    to = rng.integers(0, n_cells, 6000).astype(np.int32)
    w = rng.uniform(0.01, 1.0, 6000)
    typ = np.ones(6000, np.int32)
    lib, h = ctx.lib, ctx.h
    ip, dp = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    e, gh = C.c_void_p(), C.c_void_p()
    ctx.check(lib.cg_edges_append_host(h, C.byref(e), frm.ctypes.data_as(ip), to.ctypes.data_as(ip), w.ctypes.data_as(dp),
                                       typ.ctypes.data_as(ip), len(w)))
    ctx.check(lib.cg_assemble_graph(h, e, n_cells, C.byref(gh)))
    nv, ne = C.c_int32(0), C.c_int64(0)
    lib.cg_graph_sizes(gh, C.byref(nv), C.byref(ne))
    f = np.empty(ne.value, np.int32); t = np.empty(ne.value, np.int32); ww = np.empty(ne.value)
    ctx.check(lib.cg_graph_fetch(h, gh, None, f.ctypes.data_as(ip), t.ctypes.data_as(ip), ww.ctypes.data_as(dp), None))
    m = rng.random((nv.value, 12))
    a = ops.smooth_count_matrix(f, t, ww, m, max_n_iters=8, normalize=True, ctx=ctx)
    b = ops.smooth_count_matrix(None, None, None, m, max_n_iters=8, normalize=True, ctx=ctx, graph=gh)
    assert np.array_equal(a, b)
    lab = np.where(rng.random(nv.value) < 0.1, rng.integers(0, 5, nv.value), -1).astype(np.int32)
    pa = ops.propagate_labels(f, t, ww, nv.value, lab, 5, max_n_iters=15, ctx=ctx)
    pb = ops.propagate_labels(None, None, None, nv.value, lab, 5, max_n_iters=15, ctx=ctx, graph=gh)
    assert np.array_equal(pa, pb, equal_nan=True)
    lib.cg_graph_free(gh)
    lib.cg_edges_free(e)


def test_diffusion_edge_cases(ctx, oracle):
    from conos_b200 import ops
    m = np.arange(12, dtype=np.float64).reshape(4, 3)
    # empty graph / empty matrix: unchanged ("Empty graph passed", "Empty matrix passed")
    assert np.array_equal(ops.smooth_count_matrix(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0), m, ctx=ctx), m)
    assert ops.smooth_count_matrix([0], [1], [0.5], np.zeros((4, 0)), ctx=ctx).shape == (4, 0)
    # a vertex outside the matrix is an error, not a crash
    with pytest.raises(Exception, match="outside"):
        ops.smooth_count_matrix([0, 1], [1, 9], [0.5, 0.7], m, ctx=ctx)
    with pytest.raises(ValueError, match="must match"):
        ops.smooth_count_matrix([0, 1], [1, 2], [0.5], m, ctx=ctx)
    # one edge: the weight is (w - min) / (max - min) = 0 / 0; the NaN rows of the new matrix never win the max of the
    # norm, the norm is 0 < tol and the untouched matrix comes back -- same in the oracle port
    r = ops.smooth_count_matrix([0], [1], [0.5], m, max_n_iters=3, ctx=ctx)
    o = oracle.smooth_count_matrix([0], [1], [0.5], m, max_n_iters=3)
    assert np.array_equal(r, o) and np.array_equal(r, m)
    # correctGenes / propagateLabels need a graph
    import conos_b200
    with pytest.raises(ValueError, match="buildGraph"):
        conos_b200.Conos(ctx=ctx).propagateLabels({})


def test_correct_genes_on_the_joint_graph(ctx, oracle, small_panel):
    """Conos$correctGenes (R/conclass.R:862-893): the joint cells x genes matrix in vertex order, smoothed."""
    import conos_b200
    import scipy.sparse as sp
    con = conos_b200.Conos(small_panel, ctx=ctx)
    g = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
    res = con.correctGenes(n_od_genes=40, max_iters=6)
    genes = con._corrected_genes
    assert res.shape == (g.vcount(), len(genes)) and 0 < len(genes) <= 40
    # the same matrix assembled independently: rbind over the samples, reordered to V(graph)$name
    cells = [c for s in small_panel.values() for c in s.cells]
    blocks = []
    for s in small_panel.values():
        col = {gn: t for t, gn in enumerate(s.genes)}
        blocks.append(np.asarray(sp.csc_matrix(s.counts)[:, [col[gn] for gn in genes]].todense()))
    cm = np.vstack(blocks)
    row = {c: t for t, c in enumerate(cells)}
    cm = cm[[row[c] for c in np.asarray(g.names, dtype=object)]]
    om = oracle.smooth_count_matrix(g.edge_from, g.edge_to, g.weight, cm, max_n_iters=6, diffusion_fading=10.0,
                                    diffusion_fading_const=0.5, tol=1e-3, normalize=True)
    np.testing.assert_allclose(res, om, rtol=1e-12, atol=1e-300)
    assert np.abs(res - cm).max() > 1e-3          # it did smooth
    con.close()


def test_diffusion_matches_committed_golden_vectors(ctx):
    """The device against tests/golden/diffusion_vectors.npz (the oracle's output on the joint graph of the reference's
    own fixture, committed): no oracle run involved."""
    import importlib.util
    import os
    from conos_b200 import ops
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_diffusion_vectors", os.path.join(here, "make_diffusion_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ef, et, w, n, lab, m = mod.inputs()
    vec = np.load(os.path.join(here, "diffusion_vectors.npz"))
    for fixed in (0, 1):
        for name, tol in (("default", 0.025), ("tight", 1e-6)):
            d, it, _ = ops.propagate_labels(ef, et, w, n, lab, 3, max_n_iters=100, diffusion_fading=10.0,
                                            diffusion_fading_const=0.1, tol=tol, fixed_initial_labels=bool(fixed), ctx=ctx,
                                            return_info=True)
            assert it == int(vec[f"labels_fixed{fixed}_{name}_iters"][0])
            np.testing.assert_allclose(d, vec[f"labels_fixed{fixed}_{name}"], rtol=1e-11, atol=1e-300, equal_nan=True)
    for norm in (0, 1):
        r, it, _ = ops.smooth_count_matrix(ef, et, w, m, max_n_iters=15, diffusion_fading=10.0, diffusion_fading_const=0.5,
                                           tol=1e-3, normalize=bool(norm), ctx=ctx, return_info=True)
        assert it == int(vec[f"smooth_norm{norm}_iters"][0])
        np.testing.assert_allclose(r, vec[f"smooth_norm{norm}"], rtol=1e-11, atol=1e-300)
