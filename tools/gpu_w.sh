#!/bin/bash
# K2 hand-shake dissection (measurement builds of knn_tc2.cu, see the TC2_EXP switch)
mkdir -p gpurun_out
cp conos_b200/libconosgpu.so /tmp/full.so
for v in MMA_ONLY E1 E2 E3 E4; do cp build/libconosgpu_$v.so conos_b200/libconosgpu.so; echo $v; timeout 90 python tools/knn_time.py knn_stagger 0 0 2>&1 | tail -1; done
cp /tmp/full.so conos_b200/libconosgpu.so
