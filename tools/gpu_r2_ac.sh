#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_dropin.py -q -m gpu -k "training_loop" > gpurun_out/pytest_dropin.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dropin.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error\|Mismatch\|Max rel" gpurun_out/pytest_dropin.log | head -20
