#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -X faulthandler -m pytest tests -q -m gpu -x > gpurun_out/f_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/f_pytest.log
tail -6 gpurun_out/f_pytest.log
timeout 900 python tools/profile_stages.py --config 3 --reps 1 > gpurun_out/f_prof_config3.json 2> gpurun_out/f_prof.err || tail -5 gpurun_out/f_prof.err
cat gpurun_out/f_prof_config3.json | cut -c 1-1800
timeout 900 python tools/profile_stages.py --config 2 --reps 2 2>/dev/null | tail -1 > gpurun_out/f_prof_config2.json
cat gpurun_out/f_prof_config2.json | cut -c 1-1800
