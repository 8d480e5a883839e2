#!/bin/bash
# quick visit: fp32 parity suites + cfg5 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sparse_fused.py tests/test_gpu_parity.py -q -x > gpurun_out/pytest_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_quick.log
tail -5 gpurun_out/pytest_quick.log
timeout 400 python bench.py --workload cfg5 --batch 64 --steps 2 --warmup 1 > gpurun_out/bench_cfg5_quick.json 2> gpurun_out/bench_cfg5_quick.err
cut -c1-300 gpurun_out/bench_cfg5_quick.json
