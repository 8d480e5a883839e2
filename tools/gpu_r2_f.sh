#!/bin/bash
mkdir -p gpurun_out
for prec in bf16 bf16x2; do
for v in 1; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/l.csv python bench.py --once --precision $prec --opt fwd_fused=$v > gpurun_out/ncu_list.log 2>&1
echo "variant $prec $v"; python tools/seq_times.py gpurun_out/l.csv hshift 8 12
done
done
