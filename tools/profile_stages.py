"""Per-stage wall-clock profile of one buildGraph (library stages are stream-synchronised at both ends)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import conos_b200
from conos_b200.panel import make_panel
sys.path.insert(0, ROOT); from bench import CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--n-samples", type=int, default=None)
ap.add_argument("--n-cells", type=int, default=None)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--option", action="append", default=[], help="name=value for cg_set_option")
a = ap.parse_args()
cfg = dict(CONFIGS[a.config])
if a.n_samples: cfg["n_samples"] = a.n_samples
if a.n_cells: cfg["n_cells"] = a.n_cells
ctx = conos_b200.Context(0)
for ov in a.option:
    ctx.check(ctx.lib.cg_set_option(ctx.h, ov.split('=')[0].encode(), float(ov.split('=')[1])))
panel = make_panel(a.config, cfg["n_samples"], cfg["n_cells"], cfg["n_genes"])
con = conos_b200.Conos(panel, ctx=ctx)
kw = dict(k=cfg["k"], k_self=cfg["k_self"], space=cfg["space"], ncomps=cfg["ncomps"], n_odgenes=cfg["n_genes"])
if "alignment_strength" in cfg: kw["alignment_strength"] = cfg["alignment_strength"]
con.buildGraph(fetch=False, **kw)  # warm
for rep in range(a.reps):
    con.pairs = {}
    ctx.lib.cg_panel_drop_cache(con._panel)
    ctx.lib.cg_profile(ctx.h, 1)
    t0 = time.perf_counter()
    g = con.buildGraph(fetch=False, **kw)
    wall = time.perf_counter() - t0
    buf = C.create_string_buffer(4096)
    ctx.lib.cg_profile_dump(ctx.h, buf, 4096)
    st = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in buf.value.decode().split(";") if kv}
    ctx.lib.cg_profile(ctx.h, 0)
    top = {k: v for k, v in st.items() if "." not in k}
    print(json.dumps({"config": a.config, "n_samples": cfg["n_samples"], "n_cells": cfg["n_cells"], "space": cfg["space"],
                      "wall_ms": 1e3 * wall, "lib_ms": sum(top.values()), "edges": getattr(g, "n_edges_device", None),
                      "stages_ms": st}))
