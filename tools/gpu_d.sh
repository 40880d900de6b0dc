#!/bin/bash
mkdir -p gpurun_out
SWEEP=chunk timeout 900 python tools/jaccard_probe.py > gpurun_out/d_jaccard.log 2>&1
grep jaccard gpurun_out/d_jaccard.log
for c in 0 64 32 16; do
  timeout 600 python tools/profile_stages.py --config 2 --reps 2 --option gram_chunk_slabs=$c 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk',$c,'wall',round(d['wall_ms'],1),{k:round(v,1) for k,v in d['stages_ms'].items() if k in ('reduction','reduction.gram','reduction.eig','cross_knn','projection')})"
done
