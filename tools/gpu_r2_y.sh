#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "input_gradients" > gpurun_out/pytest_tcdx.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tcdx.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error\|error" gpurun_out/pytest_tcdx.log | head -20; grep "tc-dX" gpurun_out/tc_errors.log | cut -c1-260 | head -30
