#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), ncu launch list and one full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --once > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:shift_gemm2 -s 40 -c 1 -f -o gpurun_out/gemm2 python bench.py --once > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gemm2.ncu-rep --page raw --csv > gpurun_out/gemm2.raw.csv 2>/dev/null
tail -2 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -1; cat gpurun_out/bench_ref.json gpurun_out/bench_n1.json; head -12 gpurun_out/launch_summary.txt
