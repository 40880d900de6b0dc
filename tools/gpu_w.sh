#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_knn.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -2
python tools/knn_time.py knn_stagger 0 0 2>&1 | tail -2
python bench.py --no-e2e --no-atlas --no-check --steps 8 2> gpurun_out/w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],1), d['steps_ms'], 'knn avg launch', round(d['roofline']['avg_launch_ms'],2), 'frac', round(d['roofline']['frac'],4))"
