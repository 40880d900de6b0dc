"""Device time of the exact cross-kNN (both directions) of one n x n problem for several option values.
Fixed options for all runs: OPTS=name=value,name=value."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import conos_b200
from conos_b200 import ops
n, d, k = int(os.environ.get("N", 50000)), int(os.environ.get("D", 30)), int(os.environ.get("K", 30))
opt = sys.argv[1]
values = [float(v) for v in sys.argv[2:]]
rng = np.random.default_rng(1)
cent = rng.normal(0, 3.0, (20, d))
A = cent[rng.integers(0, 20, n)] + rng.normal(0, 1, (n, d))
B = cent[rng.integers(0, 20, n)] + rng.normal(0, 1, (n, d))
ctx = conos_b200.Context(0)
for kv in os.environ.get("OPTS", "").split(","):
    if kv:
        ctx.check(ctx.lib.cg_set_option(ctx.h, kv.split("=")[0].encode(), float(kv.split("=")[1])))
ref = None
for v in values:
    ctx.check(ctx.lib.cg_set_option(ctx.h, opt.encode(), v))
    ops.crossKnn(A, B, k, ctx=ctx, both=True)
    ctx.lib.cg_knn_timing(ctx.h, 1)
    for _ in range(5):
        out = ops.crossKnn(A, B, k, ctx=ctx, both=True)
    ms, nl, c, b = C.c_double(0), C.c_uint64(0), C.c_double(0), C.c_double(0)
    ctx.lib.cg_knn_timing_read(ctx.h, C.byref(ms), C.byref(nl), C.byref(c), C.byref(b))
    ctx.lib.cg_knn_timing(ctx.h, 0)
    same = True if ref is None else all(np.array_equal(x, y) for x, y in zip(out, ref))
    ref = out if ref is None else ref
    print(json.dumps({opt: v, "ms_per_launch": ms.value / max(nl.value, 1), "launches": nl.value, "same_result": same,
                      "redo_blocks": ctx.stat("knn_redo_blocks")}), flush=True)
ctx.close()
