#!/bin/bash
# cfg3 at the default micro-batch: launch list + one full capture of the dominant kernel (same launch shape as bench.py)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv python bench.py --once > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
head -14 gpurun_out/launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:shift_gemm2 -s 40 -c 1 -f -o gpurun_out/gemm2 python bench.py --once > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gemm2.ncu-rep --page raw --csv > gpurun_out/gemm2.raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/gemm2.raw.csv | head -12
