#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sparse_fused.py -q -m gpu -k "reorder or falls_back" > gpurun_out/pytest_reorder.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_reorder.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error" gpurun_out/pytest_reorder.log | head
