#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "horner" > gpurun_out/pytest_horner.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_horner.log
tail -3 gpurun_out/pytest_horner.log
bash tools/gpu_r2_f.sh
