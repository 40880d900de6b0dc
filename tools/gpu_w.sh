#!/bin/bash
mkdir -p gpurun_out
python tools/diffusion_probe.py > gpurun_out/w_diff.log 2>&1 && cat gpurun_out/w_diff.log | tail -1 && \
ncu --set full --clock-control none --import-source on -k regex:diffuse_step -c 3 -o gpurun_out/n_diffuse -f python tools/diffusion_probe.py > gpurun_out/w_ncu.log 2>&1; tail -2 gpurun_out/w_ncu.log
