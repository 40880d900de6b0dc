"""The default (tensor-core) mode must give the same edges for a pair whatever batch, shard or GPU count the pair is
processed in (north_star: "a joint graph identical to the exact-search reference"; the reference's result does not
depend on n.cores either: R/conclass.R:224, :1074), and the process must exit cleanly."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _table(ctx, edges):
    n = int(ctx.lib.cg_edges_count(edges))
    f, t, ty = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.int32)
    w = np.empty(n)
    ctx.check(ctx.lib.cg_edges_fetch(ctx.h, edges, f.ctypes.data_as(C.POINTER(C.c_int32)),
                                     t.ctypes.data_as(C.POINTER(C.c_int32)), w.ctypes.data_as(C.POINTER(C.c_double)),
                                     ty.ctypes.data_as(C.POINTER(C.c_int32))))
    ctx.lib.cg_edges_free(edges)
    return f, t, w, ty


def _panel(n_samples=5, cells=(900, 700, 1100, 600, 800)):
    from conos_b200.panel import make_panel
    return make_panel(11, n_samples, list(cells[:n_samples]), n_genes=400, n_types=8, n_pcs=20, use_torch=False)


@pytest.mark.parametrize("space,ncomps", [("PCA", 20), ("CPCA", 12), ("CCA", 10)])
def test_pair_shards_equal_unsharded(ctx, space, ncomps):
    """Default mode: the edges of a pair are bit-identical whether it is computed with all pairs of the panel, alone,
    or in any shard -- table rows, weights included."""
    import conos_b200
    panel = _panel()
    kw = dict(k=10, k_self=0, space=space, ncomps=ncomps, n_odgenes=400, assemble=False)
    con = conos_b200.Conos(panel, ctx=ctx)
    full = _table(ctx, con.buildGraph(**kw))
    counts = list(con._last_pair_counts.values())
    assert len(counts) == 10 and sum(counts) == len(full[0]) > 0
    offs = np.concatenate([[0], np.cumsum(counts)])
    for shard in ([0, 3, 4, 9], [1, 2], [5, 6, 7, 8], [7], [9, 0]):
        con.pairs = {}
        ctx.lib.cg_panel_drop_cache(con._panel)
        got = _table(ctx, con.buildGraph(pair_shard=shard, **kw))
        want_rows = np.concatenate([np.arange(offs[p], offs[p + 1]) for p in shard])
        for col_got, col_full in zip(got, full):
            assert np.array_equal(col_got, col_full[want_rows]), (space, shard)
    con.close()


def test_graph_independent_of_batches(ctx):
    """The batches cg_pairs_edges cuts its pair list into depend on the free device memory; the graph must not:
    one pair per batch (max_batch_bytes = 1) against all pairs in one batch."""
    import conos_b200
    panel = _panel(4)
    con = conos_b200.Conos(panel, ctx=ctx)
    g0 = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=400)
    ref = (g0.vertices.copy(), g0.edge_from.copy(), g0.edge_to.copy(), g0.weight.copy())
    ctx.lib.cg_set_option(ctx.h, b"max_batch_bytes", 1.0)
    try:
        con.pairs = {}
        ctx.lib.cg_panel_drop_cache(con._panel)
        g1 = con.buildGraph(k=10, k_self=5, space="PCA", ncomps=20, n_odgenes=400)
    finally:
        ctx.lib.cg_set_option(ctx.h, b"max_batch_bytes", float(48 << 30))
    assert np.array_equal(g1.vertices, ref[0]) and np.array_equal(g1.edge_from, ref[1])
    assert np.array_equal(g1.edge_to, ref[2]) and np.array_equal(g1.weight, ref[3])
    con.close()


@pytest.mark.parametrize("space,ncomps", [("PCA", 20), ("CPCA", 12), ("CCA", 10)])
def test_graph_independent_of_eigen_cluster_width(ctx, space, ncomps):
    """The eigen factorisation kernels run one CTA per problem when a launch holds many problems and a four-CTA cluster
    per problem when it holds few (a rank of an 8-GPU run): every sum over rows is taken over the same four slices in
    the same order in both forms, so rotations and graph are bit-identical.  Forced here through the option."""
    import conos_b200
    panel = _panel(4)
    out = {}
    try:
        for width in (1, 4):
            ctx.check(ctx.lib.cg_set_option(ctx.h, b"eig_cluster", float(width)))
            con = conos_b200.Conos(panel, ctx=ctx)
            g = con.buildGraph(k=10, k_self=5, space=space, ncomps=ncomps, n_odgenes=400)
            key = "u" if space == "CCA" else "CPC"
            out[width] = (g, {k: np.array(v[key]) for k, v in con.pairs[space].items()})
            con.close()
    finally:
        ctx.lib.cg_set_option(ctx.h, b"eig_cluster", 0.0)
    for k in out[1][1]:
        assert np.array_equal(out[1][1][k], out[4][1][k]), k
    for a in ("vertices", "edge_from", "edge_to", "weight", "type"):
        assert np.array_equal(getattr(out[1][0], a), getattr(out[4][0], a)), a


def test_clean_exit_after_neighbor_matrix():
    """The interpreter must exit with status 0 with live contexts, panels and the edge table of the last
    cg_neighbor_matrix call still around (round 1: segmentation fault at teardown)."""
    code = (
        "import numpy as np, conos_b200\n"
        "from conos_b200 import ops\n"
        "from conos_b200.panel import make_panel\n"
        "ctx = conos_b200.default_context(0)\n"
        "rng = np.random.default_rng(0)\n"
        "ops.getNeighborMatrix(rng.normal(size=(300, 20)), rng.normal(size=(200, 20)), 10, ctx=ctx)\n"
        "con = conos_b200.Conos(make_panel(7, 2, [300, 250], n_genes=200, n_types=4, n_pcs=10, use_torch=False), ctx=ctx)\n"
        "g = con.buildGraph(k=5, k_self=3, ncomps=10, n_odgenes=200)\n"
        "extra = conos_b200.Context(0)\n"
        "print('edges', g.ecount())\n")
    for snippet in (code, "import conos_b200; conos_b200.default_context(0)",
                    "import conos_b200; c = conos_b200.Context(0); c.close(); c.close()"):
        r = subprocess.run([sys.executable, "-X", "faulthandler", "-c", snippet], cwd=ROOT, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, (snippet, r.returncode, r.stdout[-2000:], r.stderr[-4000:])


def test_shutdown_with_live_handles(ctx):
    """cg_shutdown while a panel and an edge table of that context are still alive; freeing them afterwards is safe."""
    import conos_b200
    c2 = conos_b200.Context(0)
    con = conos_b200.Conos(_panel(2), ctx=c2)
    edges = con.buildGraph(k=5, k_self=0, ncomps=10, n_odgenes=400, assemble=False)
    assert c2.lib.cg_edges_count(edges) > 0
    panel_handle = con._panel
    con._panel = None
    c2.close()
    c2.lib.cg_edges_free(edges)
    c2.lib.cg_panel_free(panel_handle)
    # the session context still works
    from conos_b200 import ops
    rng = np.random.default_rng(1)
    idx, _ = ops.crossKnn(rng.normal(size=(50, 8)), rng.normal(size=(60, 8)), 5, ctx=ctx)
    assert idx.shape == (50, 5)


def _run_ranks(world, script_args, timeout=900):
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + os.getpid() % 300),
           os.path.join(ROOT, "tests", "mp_graph_worker.py")] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


@pytest.mark.parametrize("space", ["PCA", "CPCA"])
def test_two_rank_graph_equals_single_rank(space):
    """2 processes, 2 GPUs, NCCL inside the library: the gathered graph on rank 0 equals the single-rank graph bit for
    bit (vertex order, edges, weights, types)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = _run_ranks(2, ["--space", space])
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-6000:])
    assert "MP_GRAPH_EQUAL" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


def test_cpp_host_sequence_and_job():
    """tests/cuda/job_test.cpp: a C++ program that links libconosgpu.so through include/conos_gpu.h only, runs the call
    sequence of the Rcpp shim's buildGraphGPU and the same work as one cg_job call (1 device, 2 when present)."""
    exe = os.path.join(ROOT, "build", "job_test")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([exe], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "JOB_TEST_OK" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])


@pytest.mark.parametrize("space,ncomps", [("PCA", 20), ("CCA", 10)])
def test_job_path_equals_context_path(ctx, space, ncomps):
    """Conos.buildGraph(n_devices=N): ONE native call (cg_job_build_graph) -- same graph and same pairs cache as the
    call-by-call path; N = 2 when the box has two GPUs."""
    import conos_b200
    import torch
    panel = _panel(4)
    kw = dict(k=10, k_self=5, space=space, ncomps=ncomps, n_odgenes=400, score_component_variance=(space == "PCA"))
    con = conos_b200.Conos(panel, ctx=ctx)
    g0 = con.buildGraph(**kw)
    ref = {a: np.array(getattr(g0, a), copy=True) for a in ("vertices", "edge_from", "edge_to", "weight", "type")}
    for nd in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        cj = conos_b200.Conos(panel, ctx=ctx)
        g1 = cj.buildGraph(n_devices=nd, **kw)
        for a, v in ref.items():
            assert np.array_equal(getattr(g1, a), v), (nd, a)
        assert set(cj.pairs[space]) == set(con.pairs[space])
        for key, red in con.pairs[space].items():
            for f in ("CPC", "u", "v", "ul"):
                if f in red and isinstance(red[f], np.ndarray):
                    assert np.array_equal(cj.pairs[space][key][f], red[f]), (nd, key, f)
        cj.close()
    con.close()
