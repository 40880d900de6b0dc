#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
by kernel name: launches, time, share and -- when the DRAM byte counters were collected -- HBM traffic and the
achieved GB/s of every kernel (per-launch times under ncu are cold-cache and serialised: compare shares).

usage: python tools/summarize_launches.py gpurun_out/launches.csv [peak_GBs] > profiles/rNN_launches_<what>.txt
"""
import csv
import re
import sys
from collections import defaultdict

UNIT = {"ns": 1.0, "nsecond": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(path, peak):
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    per = defaultdict(dict)   # launch id -> {metric: value}, name
    names = {}
    for r in csv.DictReader(lines):
        m = r.get("Metric Name")
        if m is None:
            continue
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "")
        if m == "gpu__time_duration.sum":
            v *= UNIT.get(u, 1.0)
        elif m.startswith("dram__bytes"):
            v *= BYTES.get(u, 1.0)
        per[r["ID"]][m] = v
        n = re.sub(r"\(.*$", "", r["Kernel Name"])
        names[r["ID"]] = re.sub(r"^void ", "", n)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for i, d in per.items():
        a = agg[names[i]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    lib = lambda n: n.startswith("cg::") or "unnamed>::" in n
    own = sum(a[1] for n, a in agg.items() if lib(n))
    print(f"# {path}: {len(per)} launches, {tot / 1e6:.3f} ms total (serialised, cold-cache, under ncu)")
    print(f"# this library's kernels (cg::*): {own / 1e6:.3f} ms; everything else is the torch-side synthetic panel "
          f"generator / recall check, outside the timed region")
    have_dram = any(a[2] > 0 for a in agg.values())
    hdr = f"# {'kernel':<62} {'launches':>8} {'total ms':>10} {'share(lib)':>10} {'avg us':>9}"
    if have_dram:
        hdr += f" {'HBM MB':>10} {'GB/s':>8} {'of ' + str(int(peak)):>8}"
    print(hdr)
    for n, (c, ns, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if not lib(n):
            continue
        line = f"{n[:62]:<64} {c:>8} {ns / 1e6:>10.3f} {ns / max(own, 1):>10.3f} {ns / c / 1e3:>9.1f}"
        if have_dram:
            gbs = by / ns if ns > 0 else 0.0
            line += f" {by / 1e6:>10.1f} {gbs:>8.0f} {gbs / peak:>8.3f}"
        print(line)


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 6544.3)
