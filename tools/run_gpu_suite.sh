#!/bin/bash
# GPU call A: full GPU test-suite (exit code!), smoke, bench N=1 with stage breakdown
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
nproc >> gpurun_out/a_gpu.txt
timeout 1500 python -X faulthandler -m pytest tests -q -m gpu -s > gpurun_out/a_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/a_pytest.log
tail -30 gpurun_out/a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "SMOKE_EXIT=$?" >> gpurun_out/a_smoke.log
tail -3 gpurun_out/a_smoke.log
CONOS_BENCH_STAGES=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err; echo "BENCH_EXIT=$?" >> gpurun_out/a_bench_n1.err
tail -c 3000 gpurun_out/a_bench_n1.json
grep -E "stages rank|BENCH_EXIT|Error|error" gpurun_out/a_bench_n1.err | tail -8
