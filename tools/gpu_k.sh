#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests/test_gpu_graph.py tests/test_gpu_determinism.py tests/test_gpu_fullsize.py -q -x -s > gpurun_out/k_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/k_pytest.log
grep -E "Jaccard|passed|failed|Error|PYTEST_EXIT" gpurun_out/k_pytest.log | tail -8
timeout 900 python tools/profile_stages.py --config 2 --reps 2 2>/dev/null | tail -1 > gpurun_out/k_prof_config2.json
python -c "
import json; d=json.load(open('gpurun_out/k_prof_config2.json')); print('wall', round(d['wall_ms'],1), {k: round(v,2) for k,v in d['stages_ms'].items() if k.startswith('reduction')})"
timeout 900 python tools/profile_stages.py --config 3 --reps 1 2>/dev/null | tail -1 > gpurun_out/k_prof_config3.json
python -c "
import json; d=json.load(open('gpurun_out/k_prof_config3.json')); print('cpca wall', round(d['wall_ms'],1), {k: round(v,2) for k,v in d['stages_ms'].items() if k.startswith('reduction')})"
