"""GPU parity of the exact kNN entry points (through the C ABI) against the oracle: bit-exact indices and
float32 distances, including tie order, ragged sizes, nB < k, duplicated rows, both metrics, d > 32."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def mixture(rng, n, d, k=8, spread=3.0):
    cent = rng.normal(0, spread, (k, d))
    lab = rng.integers(0, k, n)
    return cent[lab] + rng.normal(0, 1, (n, d))


@pytest.mark.parametrize("nA,nB,d,k", [(25, 34, 30, 15), (300, 1000, 30, 15), (1000, 777, 50, 30), (513, 20, 30, 30),
                                       (257, 4099, 7, 5), (2000, 3000, 64, 32), (64, 64, 1, 3)])
def test_cross_knn_angular_bit_exact(ctx, oracle, nA, nB, d, k):
    from conos_b200 import ops
    rng = np.random.default_rng(nA * 7 + nB)
    A, B = mixture(rng, nA, d), mixture(rng, nB, d)
    m = len(B[1::5])
    B[1::5] = B[0::5][:m]   # exact duplicates -> distance ties
    iab, dab, iba, dba = ops.crossKnn(A, B, k, "angular", ctx=ctx, both=True)
    oi, od = oracle.cross_knn(A, B, k, "angular")
    assert np.array_equal(iab, oi)
    assert np.array_equal(dab.view(np.uint32), od.view(np.uint32))
    oi2, od2 = oracle.cross_knn(B, A, k, "angular")
    assert np.array_equal(iba, oi2)
    assert np.array_equal(dba.view(np.uint32), od2.view(np.uint32))


@pytest.mark.parametrize("metric,k", [("L2", 15), ("angular", 40), ("L2", 64)])
def test_cross_knn_simt_paths(ctx, oracle, metric, k):
    """k > 32 and the L2 metric are served by the exact SIMT kernel."""
    from conos_b200 import ops
    rng = np.random.default_rng(11)
    A, B = mixture(rng, 400, 30), mixture(rng, 900, 30)
    idx, dist = ops.crossKnn(A, B, k, metric, ctx=ctx)
    oi, od = oracle.cross_knn(A, B, k, metric)
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))


def test_self_knn_includes_self(ctx, oracle):
    from conos_b200 import ops
    rng = np.random.default_rng(5)
    Z = mixture(rng, 1500, 20)
    idx, dist = ops.Knn(Z, 11, "angular", ctx=ctx)
    oi, od = oracle.self_knn(Z, 11, "angular")
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    assert (idx == np.arange(len(Z))[:, None]).any(axis=1).all()


def test_zero_rows_and_sortedness(ctx, oracle):
    """All-zero rows normalise to zero vectors (distance exactly 1 to everything): ties by lowest index."""
    from conos_b200 import ops
    rng = np.random.default_rng(3)
    A, B = mixture(rng, 300, 30), mixture(rng, 500, 30)
    A[::7] = 0.0
    B[::11] = 0.0
    idx, dist = ops.crossKnn(A, B, 15, "angular", ctx=ctx)
    oi, od = oracle.cross_knn(A, B, 15, "angular")
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    assert (np.diff(dist, axis=1) >= 0).all()
    assert np.array_equal(idx[0], np.arange(15))  # zero query: every reference at distance 1


def test_recall_vs_fp64(ctx, oracle):
    from conos_b200 import ops
    rng = np.random.default_rng(9)
    A, B = mixture(rng, 2000, 30), mixture(rng, 5000, 30)
    idx, _ = ops.crossKnn(A, B, 15, "angular", ctx=ctx)
    ti, _ = oracle.knn_f64(A, B, 15, "angular")
    recall = np.mean([len(set(a) & set(b)) / 15 for a, b in zip(idx, ti)])
    assert recall > 0.999


def test_large_properties(ctx):
    """Full-size property checks (no oracle): sorted rows, valid unique indices, symmetry of mutual hits."""
    from conos_b200 import ops
    rng = np.random.default_rng(21)
    A, B = mixture(rng, 20000, 30, k=20), mixture(rng, 30000, 30, k=20)
    iab, dab, iba, dba = ops.crossKnn(A, B, 30, "angular", ctx=ctx, both=True)
    assert (np.diff(dab, axis=1) >= 0).all() and (np.diff(dba, axis=1) >= 0).all()
    assert iab.min() >= 0 and iab.max() < len(B) and iba.min() >= 0 and iba.max() < len(A)
    assert all(len(set(r)) == 30 for r in iab[::500])
    # a mutual pair must carry the identical float32 distance in both directions
    for i in range(0, 20000, 997):
        for r, j in enumerate(iab[i]):
            hit = np.nonzero(iba[j] == i)[0]
            if len(hit):
                assert dab[i, r] == dba[j, hit[0]]


@pytest.mark.parametrize("d,k", [(30, 15), (32, 15), (64, 30), (50, 30)])
def test_many_items_per_cta(ctx, oracle, d, k):
    """More 256-row work items than SMs (persistent CTAs walk several items): d = 32 / 64 leave no spare
    dimension for the folded threshold and take the compare path."""
    from conos_b200 import ops
    rng = np.random.default_rng(d * 31 + k)
    A, B = mixture(rng, 40000, d), mixture(rng, 700, d)
    iab, dab, iba, dba = ops.crossKnn(A, B, k, "angular", ctx=ctx, both=True)
    oi, od = oracle.cross_knn(A, B, k, "angular")
    assert np.array_equal(iab, oi)
    assert np.array_equal(dab.view(np.uint32), od.view(np.uint32))
    oi2, od2 = oracle.cross_knn(B, A, k, "angular")
    assert np.array_equal(iba, oi2)
    assert np.array_equal(dba.view(np.uint32), od2.view(np.uint32))


def _stat(ctx, name):
    import ctypes as C
    v = C.c_double()
    ctx.check(ctx.lib.cg_get_stat(ctx.h, name.encode(), C.byref(v)))
    return v.value


@pytest.mark.parametrize("nA,nB,d,k", [(3000, 5000, 30, 30), (700, 20000, 30, 15), (5000, 1600, 32, 10),
                                       (2500, 3100, 50, 32), (1200, 2050, 64, 20), (400, 650, 12, 30)])
def test_two_phase_kernel_equals_streaming_and_oracle(ctx, oracle, nA, nB, d, k):
    """Every kNN engine (two-phase tensor-core, streaming tensor-core, SIMT) returns the oracle's lists bit for bit;
    sizes cover all three group plans (16/32-column groups, whole half tiles, several tiles per group)."""
    from conos_b200 import ops
    rng = np.random.default_rng(nA + 3 * nB + d)
    A, B = mixture(rng, nA, d), mixture(rng, nB, d)
    B[3::17] = B[2::17][:len(B[3::17])]   # exact duplicates -> ties
    oi, od = oracle.cross_knn(A, B, k, "angular")
    try:
        for engine in (0, 1, 2):
            ctx.check(ctx.lib.cg_set_option(ctx.h, b"knn_engine", float(engine)))
            idx, dist = ops.crossKnn(A, B, k, "angular", ctx=ctx)
            assert np.array_equal(idx, oi), engine
            assert np.array_equal(dist.view(np.uint32), od.view(np.uint32)), engine
    finally:
        ctx.lib.cg_set_option(ctx.h, b"knn_engine", 0.0)


def test_two_phase_overflow_blocks_are_recomputed(ctx, oracle):
    """Candidate lists that overflow (here: capacity shrunk to 4; in production: masses of tied duplicates) hand
    their 256-row block to the streaming kernel; the result does not change."""
    from conos_b200 import ops
    rng = np.random.default_rng(77)
    A, B = mixture(rng, 1500, 30), mixture(rng, 6000, 30)
    oi, od = oracle.cross_knn(A, B, 30, "angular")
    before = _stat(ctx, "knn_redo_blocks")
    try:
        ctx.check(ctx.lib.cg_set_option(ctx.h, b"knn_cand_cap", 4.0))
        idx, dist = ops.crossKnn(A, B, 30, "angular", ctx=ctx)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"knn_cand_cap", 64.0)
    assert _stat(ctx, "knn_redo_blocks") > before
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))


def test_two_phase_mass_duplicates(ctx, oracle):
    """2000 copies of one reference point tie at the k-th distance for nearby queries: the overflow path must
    return the lowest indices among the ties."""
    from conos_b200 import ops
    rng = np.random.default_rng(78)
    A, B = mixture(rng, 600, 30), mixture(rng, 5000, 30)
    B[1000:3000] = B[999]
    A[:50] = B[999] + rng.normal(0, 1e-3, (50, 30))
    before = _stat(ctx, "knn_redo_blocks")
    idx, dist = ops.crossKnn(A, B, 15, "angular", ctx=ctx)
    oi, od = oracle.cross_knn(A, B, 15, "angular")
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    assert _stat(ctx, "knn_redo_blocks") > before


def test_two_phase_no_redo_on_regular_data(ctx):
    from conos_b200 import ops
    rng = np.random.default_rng(79)
    A, B = mixture(rng, 6000, 30, k=20), mixture(rng, 9000, 30, k=20)
    before = _stat(ctx, "knn_redo_blocks")
    ops.crossKnn(A, B, 30, "angular", ctx=ctx, both=True)
    assert _stat(ctx, "knn_redo_blocks") == before


@pytest.mark.parametrize("nA,nB,d,k", [(600, 3000, 30, 100), (300, 5000, 30, 500), (257, 6000, 30, 2250),
                                       (1000, 2500, 50, 64), (700, 1800, 32, 200), (513, 4000, 64, 333)])
def test_large_k_tensor_core_kernel(ctx, oracle, nA, nB, d, k):
    """32 < k <= 3072 (alignment.strength > 0 asks for k1 = a^2 * cells neighbours): thresholds by counting passes on
    the tensor cores, bit-identical to the oracle and to the SIMT kernel, duplicates (ties) included."""
    from conos_b200 import ops
    rng = np.random.default_rng(nA + nB + k)
    A, B = mixture(rng, nA, d), mixture(rng, nB, d)
    B[5::23] = B[4::23][:len(B[5::23])]
    oi, od = oracle.cross_knn(A, B, k, "angular")
    try:
        for engine in (0, 2):
            ctx.check(ctx.lib.cg_set_option(ctx.h, b"knn_engine", float(engine)))
            idx, dist = ops.crossKnn(A, B, k, "angular", ctx=ctx)
            assert np.array_equal(idx, oi), engine
            assert np.array_equal(dist.view(np.uint32), od.view(np.uint32)), engine
    finally:
        ctx.lib.cg_set_option(ctx.h, b"knn_engine", 0.0)


def test_large_k_mass_duplicates_fall_back(ctx, oracle):
    """Thousands of copies of one reference tie at the threshold: the candidate lists overflow and the problem is
    recomputed by the exact SIMT kernel."""
    from conos_b200 import ops
    rng = np.random.default_rng(91)
    A, B = mixture(rng, 300, 30), mixture(rng, 6000, 30)
    B[1000:5000] = B[999]
    before = _stat(ctx, "knn_redo_blocks")
    idx, dist = ops.crossKnn(A, B, 100, "angular", ctx=ctx)
    oi, od = oracle.cross_knn(A, B, 100, "angular")
    assert np.array_equal(idx, oi)
    assert np.array_equal(dist.view(np.uint32), od.view(np.uint32))
    assert _stat(ctx, "knn_redo_blocks") > before


@pytest.mark.parametrize("nA,nB,d,k,metric", [(700, 1500, 100, 15, "angular"), (300, 900, 128, 11, "angular"),
                                               (400, 600, 65, 6, "L2")])
def test_wide_panels_simt_path(ctx, oracle, nA, nB, d, k, metric):
    """65..128 columns (pagoda2's default nPcs is 100; N2R::Knn takes whatever getPca returns, R/conos.R:595):
    wider than the tensor-core operand tiles, served by the exact SIMT kernel -- bit-exact like every other path."""
    from conos_b200 import ops
    rng = np.random.default_rng(d * 1000 + k)
    A, B = rng.normal(size=(nA, d)), rng.normal(size=(nB, d))
    iab, dab, iba, dba = ops.crossKnn(A, B, k, metric, ctx=ctx, both=True)
    oi, od = oracle.cross_knn(A, B, k, metric)
    assert np.array_equal(iab, oi) and np.array_equal(dab.view(np.uint32), od.view(np.uint32))
    oi, od = oracle.cross_knn(B, A, k, metric)
    assert np.array_equal(iba, oi) and np.array_equal(dba.view(np.uint32), od.view(np.uint32))
    si, sd = ops.Knn(A, k, metric, ctx=ctx)
    oi, od = oracle.self_knn(A, k, metric)
    assert np.array_equal(si, oi) and np.array_equal(sd.view(np.uint32), od.view(np.uint32))
