#!/bin/bash
mkdir -p gpurun_out
python tools/knn_time.py knn_acc16 0 1 0 1 2>&1 | tail -4
CONOS_K2_DEBUG=1 python tools/knn_time.py knn_acc16 0 1 2>&1 | grep -E "k2" | awk '!seen[$0]++' | head -4
