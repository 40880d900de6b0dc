#!/bin/bash
mkdir -p gpurun_out
CONOS_BENCH_STAGES=1 python bench.py --no-e2e --no-atlas --no-check --steps 5 2> gpurun_out/w.err > gpurun_out/w.json
grep -E "stages rank|steps rank" gpurun_out/w.err
python tools/profile_stages.py --config 2 --reps 2 2>/dev/null | tail -40
