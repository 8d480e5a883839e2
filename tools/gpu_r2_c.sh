#!/bin/bash
# round 2, visit C: Horner-form forward — TC tests + drop-in tests, then quick bench of both tensor-core modes
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 900 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "horner" > gpurun_out/pytest_horner.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_horner.log
tail -15 gpurun_out/pytest_horner.log
timeout 1800 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -15 gpurun_out/pytest_tc.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err; tail -3 gpurun_out/bench_x2.err; cat gpurun_out/bench_x2.json
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --opt fwd_fused=0 --also "" > gpurun_out/bench_x2_unfused.json 2> gpurun_out/bench_x2_unfused.err; cat gpurun_out/bench_x2_unfused.json | cut -c1-200
