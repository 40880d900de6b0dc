#!/bin/bash
mkdir -p gpurun_out
./build/gemm_probe > gpurun_out/s_gemm_probe.log 2>&1; tail -12 gpurun_out/s_gemm_probe.log
timeout 1200 python -m pytest tests/test_gpu_graph.py -q -s 2>&1 | grep -E "JACCARD|passed|failed" | tee gpurun_out/s_jaccards.log
CONOS_BENCH_STAGES=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-atlas > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
grep -E "steps rank|stages rank" gpurun_out/s_bench.err | cut -c 1-600
