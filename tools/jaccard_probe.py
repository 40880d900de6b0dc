"""Default-mode edge set of one full-size configs[1] pair against the oracle (exact eigenvectors) for several
tolerances of the tensor-core eigen solver; prints Jaccard, principal cosines and the eigen-stage time."""
import ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conos_b200, conos_oracle as oracle
from conos_b200.panel import make_panel
from conftest import to_oracle_samples
n = int(os.environ.get("CELLS", "50000"))
panel = make_panel(2, 2, n, 2000)
osamples = to_oracle_samples(oracle, panel)
names = list(panel)
red = oracle.quick_plain_pca([osamples[names[0]], osamples[names[1]]], ncomps=30, n_odgenes=2000)
op1, op2 = oracle.project_pair([osamples[names[0]], osamples[names[1]]], red["genes"], red["CPC"][:, :30])
i2, j2, _ = oracle.neighbor_matrix(op1, op2, 30)
theirs = np.unique(i2.astype(np.int64) * (1 << 32) + j2)
ctx = conos_b200.Context(0)
SWEEP = os.environ.get("SWEEP", "tol")
cases = [(1e-5, 10, 40, 16), (3e-6, 10, 40, 16), (1e-6, 10, 60, 16)] if SWEEP == "tol" else \
    [(1e-5, 10, 40, c) for c in (0, 128, 64, 32, 16, 8)]
for tol, deg, sweeps, chunk in cases:
    ctx.check(ctx.lib.cg_set_option(ctx.h, b"gram_chunk_slabs", float(chunk)))
    ctx.check(ctx.lib.cg_set_option(ctx.h, b"tc_eig_tol", tol))
    ctx.check(ctx.lib.cg_set_option(ctx.h, b"tc_eig_degree", float(deg)))
    ctx.check(ctx.lib.cg_set_option(ctx.h, b"tc_eig_max_sweeps", float(sweeps)))
    con = conos_b200.Conos(panel, ctx=ctx)
    kw = dict(k=30, k_self=0, space="PCA", ncomps=30, n_odgenes=2000, assemble=False)
    ctx.lib.cg_edges_free(con.buildGraph(**kw))
    con.pairs = {}; ctx.lib.cg_panel_drop_cache(con._panel)
    redo0 = ctx.stat("eig_redo_pairs")
    ctx.lib.cg_profile(ctx.h, 1)
    e = con.buildGraph(**kw)
    buf = C.create_string_buffer(8192); ctx.lib.cg_profile_dump(ctx.h, buf, 8192); ctx.lib.cg_profile(ctx.h, 0)
    st = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in buf.value.decode().split(";") if kv}
    m = int(ctx.lib.cg_edges_count(e))
    f, t = np.empty(m, np.int32), np.empty(m, np.int32)
    ctx.check(ctx.lib.cg_edges_fetch(ctx.h, e, f.ctypes.data_as(C.POINTER(C.c_int32)), t.ctypes.data_as(C.POINTER(C.c_int32)), None, None))
    ctx.lib.cg_edges_free(e)
    mine = np.unique(f.astype(np.int64) * (1 << 32) + (t - n))
    inter = np.intersect1d(mine, theirs, assume_unique=True).size
    rot = con.pairs["PCA"][f"{names[0]}.vs.{names[1]}"]["CPC"]
    cosm = min(np.linalg.svd(rot[:, s::2].T @ red["CPC"][:, s::2], compute_uv=False).min() for s in (0, 1))
    print(json.dumps({"tol": tol, "degree": deg, "gram_chunk_slabs": chunk, "gram_ms": st.get("reduction.gram"), "jaccard": inter / (mine.size + theirs.size - inter), "edges": int(mine.size),
                      "oracle_edges": int(theirs.size), "min_principal_cos": float(cosm),
                      "eig_ms": st.get("reduction.eig"), "outer": st.get("reduction.eig.outer_iterations"),
                      "gemms": st.get("reduction.eig.filter_gemms"), "redo": ctx.stat("eig_redo_pairs") - redo0,
                      "worst_rel_resid": st.get("reduction.eig.worst_rel_residual_x1e9", 0) * 1e-9}), flush=True)
    con.close()
ctx.close()
