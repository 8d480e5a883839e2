#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/l.csv python bench.py --once --precision bf16 --opt fwd_fused=$v > gpurun_out/ncu_list.log 2>&1
echo "variant $v"; python tools/launch_summary.py gpurun_out/l.csv | grep "hshift\|shift_gemm2"
done
