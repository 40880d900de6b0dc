#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests/test_gpu_knn.py tests/test_gpu_fullsize.py -q -x > gpurun_out/m_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/m_pytest.log
tail -4 gpurun_out/m_pytest.log
./build/knn_probe > gpurun_out/m_knn_probe.log 2>&1; tail -12 gpurun_out/m_knn_probe.log
timeout 900 python tools/profile_stages.py --config 2 --reps 2 2>/dev/null | tail -1 > gpurun_out/m_prof_config2.json
python -c "
import json; d=json.load(open('gpurun_out/m_prof_config2.json')); print('wall', round(d['wall_ms'],1), {k: round(v,2) for k,v in d['stages_ms'].items() if '.' not in k})"
CONOS_BENCH_STAGES=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err; echo "BENCH_EXIT=$?" >> gpurun_out/m_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/m_bench_n1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "mnn_edges", "graph", "knn_recall_at_k", "graph_jaccard_vs_oracle")})
print("roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "avg_launch_ms", "share_of_step", "traffic")})
print("e2e", d.get("e2e"))
PY
grep -E "steps rank|BENCH_EXIT|Error|Traceback" gpurun_out/m_bench_n1.err | tail -4
