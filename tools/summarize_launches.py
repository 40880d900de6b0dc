#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_<what>.txt
"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        unit = r.get("Metric Unit", "ns")
        v = float(r["Metric Value"].replace(",", ""))
        ns = v * {"ns": 1.0, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1.0, "s": 1e9,
                  "second": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*$", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (serialised, cold-cache, under ncu)")
    own = sum(ns for n, ns in rows if n.startswith("cg::"))
    print(f"# this library's kernels (cg::*): {own / 1e6:.3f} ms; everything else is the torch-side synthetic panel "
          f"generator, outside the timed region")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:3]:
        if n.startswith("cg::"):
            print(f"# share of the library's kernel time: {n[:60]} = {ns / own:.3f}")
    print(f"# {'kernel':<70} {'launches':>8} {'total ms':>10} {'share':>7} {'avg us':>10}")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:<72} {c:>8} {ns / 1e6:>10.3f} {ns / tot:>7.3f} {ns / c / 1e3:>10.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
