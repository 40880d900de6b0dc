"""Host round trips of ONE device-resident configs[1] buildGraph, by call site (needs the -DCG_COUNT_SYNCS build)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import conos_b200
from conos_b200.panel import make_panel
ctx = conos_b200.Context(0)
con = conos_b200.Conos(make_panel(2, 8, 50000, 2000), ctx=ctx)
kw = dict(k=30, k_self=10, space="PCA", ncomps=30, n_odgenes=2000)
n = int(os.environ.get("CALLS", "1"))
for _ in range(n):
    con.pairs = {}
    if con._panel is not None:
        ctx.lib.cg_panel_drop_cache(con._panel)
    con.buildGraph(fetch=False, **kw)
con.close()
ctx.close()
