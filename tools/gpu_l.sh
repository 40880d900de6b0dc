#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests/test_gpu_graph.py tests/test_gpu_determinism.py -q -x > gpurun_out/l_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/l_pytest.log
tail -4 gpurun_out/l_pytest.log
# reduced configs[4]: 2 x 4 samples x 25k cells, CCA, alignment.strength 0.3 (k1 = 2250)
timeout 1500 python tools/profile_stages.py --config 5 --n-samples 8 --reps 1 2> gpurun_out/l_prof5.err | tail -1 > gpurun_out/l_prof_config5_8x25k.json
python -c "
import json; d=json.load(open('gpurun_out/l_prof_config5_8x25k.json')); print('cca 8x25k wall', round(d['wall_ms'],1), 'edges', d['edges'], {k: round(v,1) for k,v in d['stages_ms'].items()})" || tail -5 gpurun_out/l_prof5.err
timeout 900 python tools/profile_stages.py --config 3 --reps 1 2>/dev/null | tail -1 > gpurun_out/l_prof_config3.json
python -c "
import json; d=json.load(open('gpurun_out/l_prof_config3.json')); print('cpca wall', round(d['wall_ms'],1), {k: round(v,2) for k,v in d['stages_ms'].items() if k.startswith('reduction')})"
