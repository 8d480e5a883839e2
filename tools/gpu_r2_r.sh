#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sparse_fused.py -q -m gpu -k "cfg5_generator" > gpurun_out/pytest_kink.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_kink.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error" gpurun_out/pytest_kink.log | head; cat gpurun_out/cfg5_kink_proof.json
