#!/bin/bash
# sparse (cfg5) visit: fused-kernel parity tests, cfg5 bench, launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sparse_fused.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_sparse.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sparse.log
tail -15 gpurun_out/pytest_sparse.log
timeout 600 python bench.py --workload cfg5 --batch 64 --steps 2 --warmup 1 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
cat gpurun_out/bench_cfg5.json; tail -5 gpurun_out/bench_cfg5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --once > gpurun_out/ncu_list_cfg5.log 2>&1
python tools/launch_summary.py gpurun_out/launches_cfg5.csv > gpurun_out/launch_summary_cfg5.txt 2>&1
cat gpurun_out/launch_summary_cfg5.txt
