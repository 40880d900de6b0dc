#!/bin/bash
# full verification: GPU suite, then the default bench line
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/u_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/u_tests.log
tail -5 gpurun_out/u_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/u_smoke.log
python bench.py > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err; echo "bench rc=$?"
cat gpurun_out/u_bench.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/u_ref.json 2> gpurun_out/u_ref.err; echo "ref rc=$?"
cat gpurun_out/u_ref.json
