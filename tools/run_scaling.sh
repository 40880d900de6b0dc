#!/bin/bash
# one multi-GPU bench line: bash tools/run_scaling.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/s_bench_n$N.json 2> gpurun_out/s_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/s_bench_n$N.json"))
print("N=$N", "ms", round(d["ms_per_step"],2), d["steps_ms"], "value", round(d["value"],1), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step_median"], "atlas", d["atlas"]["ms_per_step"], "same graph", d.get("graph_equals_single_rank"))
PY
