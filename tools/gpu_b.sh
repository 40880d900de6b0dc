#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_determinism.py -q -s > gpurun_out/b_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/b_pytest.log
tail -5 gpurun_out/b_pytest.log
CONOS_BENCH_STAGES=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench_n1.json 2> gpurun_out/b_bench_n1.err; echo "BENCH_EXIT=$?" >> gpurun_out/b_bench_n1.err
tail -c 2500 gpurun_out/b_bench_n1.json
grep -E "stages rank|BENCH_EXIT|Error|error" gpurun_out/b_bench_n1.err | tail -8
timeout 900 python tools/jaccard_probe.py > gpurun_out/b_jaccard.log 2>&1
grep jaccard gpurun_out/b_jaccard.log
