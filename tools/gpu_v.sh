#!/bin/bash
# what runs beside the bench when steps are slow?
mkdir -p gpurun_out
( for i in $(seq 1 150); do date +%s.%N; ps -eo pid,ppid,comm,pcpu,etimes --no-headers | grep -v "\[" | grep -v -E " (ps|grep|sleep|date|seq)$| (ps|grep|sleep|date|seq) " ; grep "^cpu " /proc/stat; sleep 0.1; done ) > gpurun_out/v_watch.txt 2>&1 &
W=$!
CONOS_BENCH_NO_SAMPLER=1 CONOS_BENCH_NO_KNN_TIMING=1 python bench.py --no-e2e --no-check --no-atlas --steps 120 2> gpurun_out/v.err | python -c "
import json,sys,time
d=json.loads(sys.stdin.read()); s=d['steps_ms']; import itertools
t=list(itertools.accumulate(s)); med=sorted(s)[len(s)//2]
print('end', time.time(), 'mean', round(d['ms_per_step'],1), 'median', med, 'outliers (t_ms, step_ms):', [(round(t[i]-s[i]), s[i]) for i in range(len(s)) if s[i] > med+4], 'total', t[-1])"
kill $W 2>/dev/null
grep -c . gpurun_out/v_watch.txt
