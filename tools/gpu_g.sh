#!/bin/bash
# 2-GPU box: Gram sharing parity + bench
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_determinism.py -q -k two_rank -s > gpurun_out/g_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/g_pytest.log
tail -8 gpurun_out/g_pytest.log
CONOS_BENCH_STAGES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g_bench_n2.json 2> gpurun_out/g_bench_n2.err; echo "BENCH_EXIT=$?" >> gpurun_out/g_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/g_bench_n2.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph", "graph_equals_single_rank", "ranks")})
    print("e2e", d.get("e2e"))
except Exception as e:
    print("no bench line", e)
PY
grep -E "stages rank|BENCH_EXIT|Error|error|Traceback" gpurun_out/g_bench_n2.err | tail -12
