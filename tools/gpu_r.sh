#!/bin/bash
mkdir -p gpurun_out
python tools/knn_time.py knn_ld32 0 1 0 1 > gpurun_out/r_ld32.log 2>&1; cat gpurun_out/r_ld32.log | tail -5
