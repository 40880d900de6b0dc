"""Where the host side of one atlas buildGraph (configs[3]: 100 samples x 10k cells, 4950 pairs) spends its time."""
import cProfile, json, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import conos_b200
from conos_b200.panel import make_panel
ctx = conos_b200.Context(0)
panel = make_panel(4, 100, 10000, 2000)
con = conos_b200.Conos(panel, ctx=ctx)
kw = dict(k=15, k_self=5, space="PCA", ncomps=30, n_odgenes=2000)
con.buildGraph(fetch=False, **kw)
for rep in range(2):
    con.pairs = {}
    ctx.lib.cg_panel_drop_cache(con._panel)
    t0 = time.perf_counter()
    if rep == 1:
        pr = cProfile.Profile()
        pr.enable()
    con.buildGraph(fetch=False, **kw)
    if rep == 1:
        pr.disable()
    print(json.dumps({"wall_ms": round(1e3 * (time.perf_counter() - t0), 1), **{k: round(v, 1) for k, v in con.timings.items()}}))
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(18)
con.close()
ctx.close()
