#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (read here, no GPU): python tools/ncu_summary.py file.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.avg",
        "sm__inst_executed.sum.per_cycle_active", "launch__cluster_size", "launch__grid_size", "launch__block_size",
        "smsp__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]


def main(paths):
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"# {path}: no data")
            continue
        hdr, units = rows[0], rows[1]
        print(f"# {path}")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"\nKernel = {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for i, h in enumerate(hdr):
                if any(h.endswith(k) or h == k for k in KEEP):
                    print(f"  {h} [{units[i]}] = {r[i]}")


if __name__ == "__main__":
    main(sys.argv[1:])
