#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -X faulthandler -m pytest tests -q -m gpu > gpurun_out/p_pytest.log 2>&1; echo "PYTEST_EXIT=$?" >> gpurun_out/p_pytest.log
grep -E "^FAILED|^ERROR|Error:|passed|failed|PYTEST_EXIT" gpurun_out/p_pytest.log | tail -12
