"""propagate_labels on a synthetic graph with the size and locality of the configs[1] joint graph (8 samples x 50k
cells, ~31 neighbours per cell, most of them inside a band of the vertex order) -- the launch ncu captures."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import conos_b200
from conos_b200 import ops
n, deg, n_lab = 400_000, 16, int(os.environ.get("LABELS", 20))
rng = np.random.default_rng(0)
a = np.repeat(np.arange(n, dtype=np.int64), deg)
b = (a + rng.integers(1, 50_000, len(a)) * rng.choice([1, 8], len(a))) % n      # same sample or another one
lo, hi = np.minimum(a, b), np.maximum(a, b)
key = np.unique(lo * n + hi)
ef, et = (key // n).astype(np.int32), (key % n).astype(np.int32)
w = rng.uniform(0.05, 1.0, len(ef))
lab = np.where(rng.random(n) < 0.1, rng.integers(0, n_lab, n), -1).astype(np.int32)
ctx = conos_b200.Context(0)
ops.propagate_labels(ef, et, w, n, lab, n_lab, max_n_iters=1, tol=0.0, fixed_initial_labels=True, ctx=ctx)
ts = {}
for iters in (2, 22):
    t0 = time.perf_counter()
    ops.propagate_labels(ef, et, w, n, lab, n_lab, max_n_iters=iters, tol=0.0, fixed_initial_labels=True, ctx=ctx)
    ts[iters] = time.perf_counter() - t0
print(json.dumps({"vertices": n, "edges": len(ef), "labels": n_lab, "ms_per_iteration": 1e3 * (ts[22] - ts[2]) / 20}))
ctx.close()
