#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -X faulthandler -m pytest tests -q -m gpu > gpurun_out/o_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/o_pytest.log
tail -6 gpurun_out/o_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/o_bench_n1.json 2> gpurun_out/o_bench_n1.err; echo "BENCH_EXIT=$?" >> gpurun_out/o_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/o_bench_n1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph", "knn_recall_at_k", "graph_jaccard_vs_oracle", "gpu_launches")})
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step", "traffic")})
print("e2e", d.get("e2e")); print("cpu", d.get("cpu_baseline"))
PY
grep -E "BENCH_EXIT|Error|Traceback" gpurun_out/o_bench_n1.err | tail -3
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/o_bench_ref.json 2> gpurun_out/o_bench_ref.err; cut -c 1-700 gpurun_out/o_bench_ref.json
timeout 1200 python bench.py --config 4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/o_bench_atlas_n1.json 2> gpurun_out/o_bench_atlas_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/o_bench_atlas_n1.json").read().strip().splitlines()[-1])
print("atlas n1", {k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph")}, "e2e", d.get("e2e"))
PY
