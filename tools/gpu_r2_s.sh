#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "node_gated or auto_precision" > gpurun_out/pytest_tcnode.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tcnode.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error\|error" gpurun_out/pytest_tcnode.log | head -20; grep "tc-node" gpurun_out/tc_errors.log | head -30
