#!/usr/bin/env python
"""Hottest source lines of one kernel from an ncu report: python tools/ncu_hot_lines.py report.ncu-rep [N]
(needs the capture to be made with --import-source on and the code compiled with -lineinfo)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main(rep, top=25):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    for k, r in enumerate(rows):
        if "Source" in r and "Instructions Executed" in r:
            hdr, body = r, rows[k + 1:]
            break
    if hdr is None:
        print(out[:2000])
        return
    isrc = hdr.index("Source")
    ist = hdr.index("Warp Stall Sampling (All Samples)")
    ie = hdr.index("Instructions Executed")
    tot_s = tot_e = 0
    agg = defaultdict(lambda: [0, 0])
    for r in body:
        if len(r) <= max(ist, ie):
            continue
        try:
            s, e = int(r[ist] or 0), int(r[ie] or 0)
        except ValueError:
            continue
        agg[r[isrc][:110]][0] += s
        agg[r[isrc][:110]][1] += e
        tot_s += s
        tot_e += e
    print(f"# {rep}: {tot_s} stall samples, {tot_e} warp instructions")
    for src, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{s / max(tot_s, 1):6.3f} {e / max(tot_e, 1):6.3f}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
