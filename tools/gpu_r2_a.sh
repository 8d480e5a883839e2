#!/bin/bash
# round 2, visit A: split-bf16 kernels — unit tests, full-horizon calibration, quick bench of both tensor-core modes
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1200 python -m pytest tests/test_gpu_tc.py -q -m gpu -x --deselect tests/test_gpu_tc.py::test_tc_full_size_batch_properties > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -25 gpurun_out/pytest_tc.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err; tail -3 gpurun_out/bench_x2.err; cat gpurun_out/bench_x2.json
