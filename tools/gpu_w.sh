#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
python bench.py --no-atlas --no-check --steps 150 2> gpurun_out/w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['steps_ms']; m=sorted(s)[len(s)//2]; print('mean', round(d['ms_per_step'],1), 'median', m, 'n>median+4:', sum(1 for x in s if x>m+4), 'max', max(s), 'e2e', d['e2e'])"
