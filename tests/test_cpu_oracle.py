"""The oracle against independent restatements and its own invariants (CPU only).  The reference ships no
golden vectors for this path (parity unpinned); these tests pin the oracle's pieces against
slow, obviously-correct NumPy / pure-Python versions and against the committed regression fixtures."""
import os

import math

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fma32(a, b, c):
    """float32 fma through float64 (exact product; the one float64 rounding of the sum is a double rounding
    only in ~2^-29 of the cases)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def slow_knn(q, r, k, metric):
    nq, d = q.shape
    acc = [np.zeros((nq, r.shape[0]), np.float32) for _ in range(4)]
    for c in range(d):
        if metric == "angular":
            acc[c & 3] = fma32(np.broadcast_to(q[:, c:c + 1], acc[0].shape), np.broadcast_to(r[None, :, c], acc[0].shape),
                               acc[c & 3])
        else:
            df = (q[:, c:c + 1] - r[None, :, c]).astype(np.float32)
            acc[c & 3] = fma32(df, df, acc[c & 3])
    tot = (acc[0] + acc[1]) + (acc[2] + acc[3])
    dist = (np.float32(1.0) - tot) if metric == "angular" else tot
    idx = np.argsort(dist, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int32), np.take_along_axis(dist, idx, axis=1)


@pytest.mark.parametrize("metric", ["angular", "L2"])
@pytest.mark.parametrize("nq,nr,d,k", [(40, 60, 30, 15), (33, 17, 7, 5), (50, 200, 50, 30)])
def test_knn_f32_canonical_arithmetic(oracle, metric, nq, nr, d, k):
    rng = np.random.default_rng(nq + nr)
    A, B = rng.normal(size=(nq, d)), rng.normal(size=(nr, d))
    B[3] = B[1]
    q, r = oracle.prepare_f32(A, metric), oracle.prepare_f32(B, metric)
    idx, dist = oracle.knn_f32(q, r, k, metric)
    si, sd = slow_knn(q, r, min(k, nr), metric)
    assert np.array_equal(idx[:, :si.shape[1]], si)
    assert np.array_equal(dist[:, :si.shape[1]].view(np.uint32), sd.view(np.uint32))
    if metric == "angular":
        nrm = np.sqrt((q.astype(np.float64) ** 2).sum(axis=1))
        assert np.allclose(nrm, 1.0, atol=1e-6)


def test_knn_fewer_refs_than_k(oracle):
    rng = np.random.default_rng(0)
    idx, dist = oracle.cross_knn(rng.normal(size=(5, 4)), rng.normal(size=(3, 4)), 6)
    assert (idx[:, 3:] == -1).all() and np.isinf(dist[:, 3:]).all() and (idx[:, :3] >= 0).all()


def test_knn_f32_close_to_f64(oracle):
    rng = np.random.default_rng(1)
    A, B = rng.normal(size=(300, 30)), rng.normal(size=(500, 30))
    i32, d32 = oracle.cross_knn(A, B, 10)
    i64, d64 = oracle.knn_f64(A, B, 10)
    assert np.mean(i32 == i64) > 0.995
    assert np.abs(d32 - d64).max() < 1e-6


def test_pare_down_hub_edges_against_python(oracle):
    rng = np.random.default_rng(2)
    m = sp.random(60, 40, density=0.3, random_state=3, format="csc")
    m.sort_indices()
    k, klow = 5, 3
    rowN = np.bincount(m.indices, minlength=60).astype(np.int32)
    want_x = m.data.copy()
    want_rowN = rowN.copy()
    for g in range(m.shape[1]):                       # src/edgeFilter.cpp:33-55, stable descending sort
        p0, p1 = m.indptr[g], m.indptr[g + 1]
        ne = p1 - p0
        if ne <= k:
            continue
        cid = want_rowN[m.indices[p0:p1]].copy()
        order = np.argsort(-cid, kind="stable")
        j = 0
        while ne > k and j < len(order) and cid[order[j]] > klow:
            want_x[order[j] + p0] = 0
            want_rowN[m.indices[order[j] + p0]] -= 1
            ne -= 1
            j += 1
    got = oracle.pare_down_hub_edges(m, rowN, k, klow)
    assert np.array_equal(got, want_x) and np.array_equal(rowN, want_rowN)


def test_reduce_edges_degree_bounds(oracle):
    adj = sp.random(200, 180, density=0.2, random_state=5, format="csc")
    out = oracle.reduce_edges_in_graph_iteratively(adj, 10)
    cc = np.diff(out.indptr)
    rc = np.bincount(out.indices, minlength=200)
    assert max(cc.max(), rc.max()) <= 10 + 5          # max.kdiff tolerance of the reference
    assert out.nnz < adj.nnz
    # surviving entries are a subset with unchanged values
    d = (adj - out).tocsr()
    assert (abs(out.multiply(out != 0) - adj.multiply(out != 0))).max() == 0 and d.min() >= 0


def test_cpca_against_numpy(oracle):
    rng = np.random.default_rng(4)
    p, k, ncomp = 12, 2, 4
    covs = []
    for _ in range(k):
        a = rng.normal(size=(40, p))
        covs.append(np.cov(a, rowvar=False))
    ng = [40.0, 55.0]
    S = sum(n / sum(ng) * c for n, c in zip(ng, covs))
    w, v = np.linalg.eigh(S)
    ev = v[:, ::-1][:, :ncomp]
    got = oracle.cpca_f(covs, ng, ncomp, 500, 1e-5, ev)
    # plain NumPy restatement of src/cpca.cpp:44-91
    Qw = np.eye(p)
    CPC = np.zeros((p, ncomp))
    D = np.zeros((ncomp, k))
    nit = np.zeros(ncomp)
    for comp in range(ncomp):
        q = ev[:, comp].copy()
        d = np.array([q @ (c @ q) for c in covs])
        cost0 = 0.0
        it = 0
        for it in range(500):
            Sm = sum(c * (n / di) for c, n, di in zip(covs, ng, d))
            wv = Sm @ q
            if comp != 0:
                wv = Qw @ wv
            q = wv / np.sqrt(np.sum(wv * wv))
            d = np.array([q @ (c @ q) for c in covs])
            cost = float(np.sum(np.log(d) * np.array(ng)))
            if abs((cost - cost0) / cost) < 1e-5:
                break
            cost0 = cost
        else:
            it = 500
        nit[comp] = it
        D[comp] = d
        CPC[:, comp] = q
        Qw -= np.outer(q, q)
    assert np.array_equal(got["niter"], nit)
    np.testing.assert_allclose(got["CPC"], CPC, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got["D"], D, rtol=1e-10)


def test_assemble_graph_small_example(oracle):
    # rows: (5,3,w) (3,5,w) duplicate undirected; (2,2) loop; (7,1,0) dropped before numbering
    frm = np.array([5, 3, 2, 7, 5, 9])
    to = np.array([3, 5, 2, 1, 9, 3])
    w = np.array([0.5, 0.25, 1.0, 0.0, 0.125, 2.0])
    typ = np.array([1, 0, 0, 1, 1, 0])
    g = oracle.assemble_graph(frm, to, w, typ)
    assert g["vertices"].tolist() == [5, 3, 2, 9]          # first appearance; 7 and 1 never appear
    assert list(zip(g["from"], g["to"])) == [(0, 1), (0, 3), (1, 3)]
    assert g["weight"].tolist() == [0.75, 0.125, 2.0] and g["type"].tolist() == [1, 1, 0]


def test_scalars(oracle):
    assert oracle.r_round(0.5) == 0 and oracle.r_round(1.5) == 2 and oracle.r_round(2.5) == 2
    assert oracle.resolve_k1(15, None, None, 1000) == (15, 0.0)
    assert oracle.resolve_k1(15, None, 0.3, 25000) == (2250, 0.3)
    assert oracle.resolve_k1(15, None, 5.0, 100)[0] == 100
    assert oracle.cor_base_of(0.0) == 1.0 and oracle.cor_base_of(0.05) == 1.5 and oracle.cor_base_of(0.3) == 2.0
    with pytest.raises(ValueError):
        oracle.resolve_k1(15, 10, None, 100)


def test_centering_quirk(oracle, small_panel):
    """common.centering=TRUE subtracts ONE NUMBER per sample (m[1], m[2]) -- bug-compatible default."""
    from conftest import to_oracle_samples
    s = list(to_oracle_samples(oracle, small_panel).values())[:2]
    od = oracle.common_overdispersed_genes(s, 300)
    rot = np.linalg.qr(np.random.default_rng(0).normal(size=(len(od), 6)))[0]
    p = oracle.project_pair(s, od, rot, True, True)
    sm = oracle.scaled_matrices(s, od, True)
    n = np.array([x.shape[0] for x in sm], float)
    means = np.stack([np.asarray(x.mean(axis=0)).ravel() for x in sm])
    m = (means * n[:, None]).sum(0) / n.sum()
    for side in (0, 1):
        want = sm[side] @ rot - m[side] * rot.sum(axis=0)[None, :]
        np.testing.assert_allclose(p[side], want, rtol=1e-10, atol=1e-12)


def test_build_graph_invariants(oracle, small_panel):
    from conftest import to_oracle_samples
    S = to_oracle_samples(oracle, small_panel)
    r = oracle.build_graph(S, k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=300)
    g = r["graph"]
    assert (g["from"] < g["to"]).all()
    key = g["from"].astype(np.int64) * len(g["vertices"]) + g["to"]
    assert (np.diff(key) > 0).all()                       # sorted, unique
    assert (g["weight"] > 0).all()
    offs = np.cumsum([0] + [s.counts.shape[0] for s in S.values()])
    sample_of = np.searchsorted(offs, g["vertices"], side="right") - 1
    inter = sample_of[g["from"]] != sample_of[g["to"]]
    assert np.array_equal(inter, g["type"] == 1)
    assert g["weight"][inter].max() <= 1.0 + 1e-12 and g["weight"][~inter].max() <= 0.2 + 1e-12
    a = oracle.adjacency_csr(g)
    assert abs(a - a.T).max() == 0


def test_reference_fixture_and_golden_vectors(oracle, golden):
    """The reference's own panel (data/small_panel.preprocessed.rda, decoded without R) through the oracle:
    the three quantities tests/testthat/test_functions.R asserts on this path -- 2 samples (:10), 59 cells (:15),
    59 x 1000 joint count matrix (:20), and one graph vertex per cell after buildGraph(ncomps=25) (:24-33:
    length(groups) == 59) -- and the oracle's committed output on it, so the oracle cannot drift silently."""
    mod, vec = golden
    panel = mod.load_reference_panel(oracle.Sample)
    assert len(panel) == 2
    assert sum(s.counts.shape[0] for s in panel.values()) == 59
    genes = sorted(set(g for s in panel.values() for g in s.genes))
    assert 59 * len(genes) == 59000
    ref = oracle.build_graph(panel, ncomps=25)
    g = ref["graph"]
    assert len(g["vertices"]) == 59
    assert np.array_equal(g["vertices"], vec["graph_vertices"])
    assert np.array_equal(g["from"], vec["graph_from"]) and np.array_equal(g["to"], vec["graph_to"])
    assert np.array_equal(g["type"], vec["graph_type"])
    np.testing.assert_allclose(g["weight"], vec["graph_weight"], rtol=1e-10)
    g0 = oracle.build_graph(panel, k=15, k_self=5, space="PCA", ncomps=30)["graph"]   # BASELINE.json configs[0]
    assert len(g0["vertices"]) == 59
    assert np.array_equal(g0["from"], vec["c0_from"]) and np.array_equal(g0["to"], vec["c0_to"])
    np.testing.assert_allclose(g0["weight"], vec["c0_weight"], rtol=1e-10)
    A, B = mod.knn_inputs()
    idx, dist = oracle.cross_knn(A, B, 15, "angular")
    assert np.array_equal(idx, vec["knn_idx"])
    assert np.array_equal(dist.view(np.uint32), vec["knn_dist"].view(np.uint32))


def test_smooth_count_matrix_against_slow_restatement(oracle):
    """src/propagate_labels.cpp:120-175 restated edge by edge in NumPy (rows as vectors, the reference's own loop order)
    against the C oracle: bit-identical, including the stop-before-update rule and fixed vertices."""
    rng = np.random.default_rng(11)
    n, n_edges, n_cols = 60, 400, 5
    ef = rng.integers(0, n, n_edges)
    et = rng.integers(0, n, n_edges)
    w = rng.uniform(0.1, 1.0, n_edges)
    m0 = rng.random((n, n_cols))
    fixed = rng.random(n) < 0.25

    def slow(max_iters, fading, const, tol, normalize, is_fixed):
        cm = m0.copy()
        min_w = max(w.min(), 0.0)
        max_w = w.max() - min_w
        length = 1.0 - (w - min_w) / max_w
        ew = np.array([math.exp(-fading * (l + const)) for l in length])   # libm, like the C oracle (np.exp is SIMD)
        sw = np.ones(n)
        it, nrm = 0, 1e10
        for it in range(max_iters):
            new = cm.copy()
            for e in range(n_edges):
                a, b = ef[e], et[e]
                if is_fixed is None or not is_fixed[a]:
                    new[a] = new[a] + cm[b] * ew[e]
                    if normalize:
                        sw[a] += ew[e]
                if is_fixed is None or not is_fixed[b]:
                    new[b] = new[b] + cm[a] * ew[e]
                    if normalize:
                        sw[b] += ew[e]
            if normalize:
                new = new / sw[:, None]
                sw[:] = 1.0
            nrm = np.abs(new - cm).max()
            if nrm < tol:
                break
            cm = new
        else:
            it = max_iters
        return cm, it, nrm

    for args in ((4, 1.0, 0.1, 1e-3, False, None), (6, 10.0, 0.5, 5e-3, True, fixed), (40, 2.0, 0.1, 0.02, True, None)):
        sm, sit, snrm = slow(*args)
        om, oit, onrm = oracle.smooth_count_matrix(ef, et, w, m0, args[5], max_n_iters=args[0], diffusion_fading=args[1],
                                                   diffusion_fading_const=args[2], tol=args[3], normalize=args[4],
                                                   return_info=True)
        assert oit == sit
        assert np.array_equal(om, sm)
        assert onrm == snrm


def _diffusion_golden():
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_diffusion_vectors", os.path.join(here, "make_diffusion_vectors.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.inputs(), np.load(os.path.join(here, "diffusion_vectors.npz"))


def test_diffusion_golden_vectors(oracle):
    """The committed diffusion vectors (tests/golden/make_diffusion_vectors.py: the fixture's joint graph) are what the
    oracle gives today: iteration counts equal, values to 1e-13 (exp() of another libm may differ in the last bit)."""
    (ef, et, w, n, lab, m), vec = _diffusion_golden()
    for fixed in (0, 1):
        for name, tol in (("default", 0.025), ("tight", 1e-6)):
            d, it, _ = oracle.propagate_labels(ef, et, w, n, lab, 3, max_n_iters=100, diffusion_fading=10.0,
                                               diffusion_fading_const=0.1, tol=tol, fixed_initial_labels=bool(fixed),
                                               return_info=True)
            assert it == int(vec[f"labels_fixed{fixed}_{name}_iters"][0])
            np.testing.assert_allclose(d, vec[f"labels_fixed{fixed}_{name}"], rtol=1e-13, atol=1e-300, equal_nan=True)
            if fixed:
                assert np.array_equal(d[lab >= 0], np.eye(3)[lab[lab >= 0]])   # fixed labels stay one-hot
    for norm in (0, 1):
        r, it, _ = oracle.smooth_count_matrix(ef, et, w, m, max_n_iters=15, diffusion_fading=10.0,
                                              diffusion_fading_const=0.5, tol=1e-3, normalize=bool(norm), return_info=True)
        assert it == int(vec[f"smooth_norm{norm}_iters"][0])
        np.testing.assert_allclose(r, vec[f"smooth_norm{norm}"], rtol=1e-13, atol=1e-300)
