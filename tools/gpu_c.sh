#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -q -m gpu -x > gpurun_out/c_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/c_pytest.log
tail -25 gpurun_out/c_pytest.log
timeout 900 python tools/jaccard_probe.py > gpurun_out/c_jaccard.log 2>&1
grep jaccard gpurun_out/c_jaccard.log | head -3
for cfg in "--config 3 --n-samples 8" "--config 3"; do
  timeout 900 python tools/profile_stages.py $cfg --reps 1 > gpurun_out/c_prof_$(echo $cfg | tr -d ' -').json 2> gpurun_out/c_prof.err || tail -5 gpurun_out/c_prof.err
done
cat gpurun_out/c_prof_*.json | cut -c 1-1500
CONOS_BENCH_STAGES=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err; echo "BENCH_EXIT=$?" >> gpurun_out/c_bench_n1.err
cut -c 1-600 gpurun_out/c_bench_n1.json
grep -E "stages rank|BENCH_EXIT|Error|error" gpurun_out/c_bench_n1.err | tail -8
