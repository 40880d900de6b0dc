#!/bin/bash
# 2-GPU box: full determinism suite (cg_job, C++ test, NCCL packing) + bench N=2
mkdir -p gpurun_out
timeout 900 ./build/job_test > gpurun_out/i_jobtest.log 2>&1; echo "JOBTEST_EXIT=$?" >> gpurun_out/i_jobtest.log; tail -5 gpurun_out/i_jobtest.log
timeout 1200 python -X faulthandler -m pytest tests/test_gpu_determinism.py -q -x > gpurun_out/i_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/i_pytest.log
tail -15 gpurun_out/i_pytest.log
CONOS_BENCH_STAGES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/i_bench_n2.json 2> gpurun_out/i_bench_n2.err; echo "BENCH_EXIT=$?" >> gpurun_out/i_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/i_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph", "graph_equals_single_rank")})
    print("e2e", d.get("e2e"))
except Exception as e:
    print("no bench line", e)
PY
grep -E "stages rank|steps rank|BENCH_EXIT|Error|error|Traceback" gpurun_out/i_bench_n2.err | tail -8
