#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_graph.py tests/test_gpu_determinism.py -q -m gpu -x 2>&1 | tail -3
python tools/atlas_py_profile.py 2>&1 | grep wall_ms
python tools/profile_stages.py --config 4 --reps 1 2>/dev/null | tail -1
