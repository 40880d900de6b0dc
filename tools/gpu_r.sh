#!/bin/bash
mkdir -p gpurun_out
python tools/knn_time.py knn_stagger 0 200 400 600 800 1200 0 > gpurun_out/r_stagger.log 2>&1; cat gpurun_out/r_stagger.log | tail -8
