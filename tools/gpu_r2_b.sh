#!/bin/bash
# round 2, visit B: whole GPU suite, smoke, x2 bench with CPU baseline, launch list of one x2 micro-batch
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_x2.csv python bench.py --once > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches_x2.csv > gpurun_out/launch_summary_x2.txt 2>&1; head -16 gpurun_out/launch_summary_x2.txt
