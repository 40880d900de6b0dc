#!/bin/bash
# 8-GPU confirmation: multi-GPU tests, then the bench line (configs[1] + atlas sub-record)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_determinism.py -q -m gpu 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/x_bench_n8.json 2> gpurun_out/x_bench_n8.err; echo "rc=$?"
cat gpurun_out/x_bench_n8.json
