#!/bin/bash
# 8-GPU box: configs[1] at N=8 and N=4, atlas (configs[3]) at N=8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/h_gpus.txt
run() { # tag nproc args...
  tag=$1; np=$2; shift 2
  CONOS_BENCH_STAGES=1 timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/h_$tag.json 2> gpurun_out/h_$tag.err; echo "BENCH_EXIT=$?" >> gpurun_out/h_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/h_{tag}.json").read().strip().splitlines()[-1])
    print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph", "graph_equals_single_rank")})
    print(tag, "ranks", d.get("ranks"))
    print(tag, "e2e", d.get("e2e"))
except Exception as e:
    print(tag, "no bench line", e)
PY
  grep -E "stages rank 0|stages rank 1\]|BENCH_EXIT|Error|Traceback" gpurun_out/h_$tag.err | tail -4
}
run c2_n8 8 --steps 5 --warmup 3
run c2_n4 4 --steps 5 --warmup 3
run atlas_n8 8 --config 4 --steps 2 --warmup 1
