#!/bin/bash
# 8-GPU box: the bench line as the driver runs it (atlas sub-record included) at N=8 and N=4
mkdir -p gpurun_out
run() { tag=$1; np=$2; shift 2
  CONOS_BENCH_STAGES=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $np "$@" > gpurun_out/q_$tag.json 2> gpurun_out/q_$tag.err; echo "BENCH_EXIT=$?" >> gpurun_out/q_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/q_{tag}.json").read().strip().splitlines()[-1])
    print(tag, {k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph", "graph_equals_single_rank")})
    print(tag, "e2e", d.get("e2e"))
    print(tag, "atlas", d.get("atlas"))
except Exception as e:
    print(tag, "no bench line", e)
PY
  grep -E "stages rank 0|steps rank 0|BENCH_EXIT|Error|Traceback" gpurun_out/q_$tag.err | tail -4
}
run c2_n8 8 --steps 5 --warmup 3
run c2_n4 4 --steps 5 --warmup 3
