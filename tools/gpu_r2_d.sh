#!/bin/bash
# launch lists of one micro-batch (forward part is enough): Horner forward, x2 and bf16
mkdir -p gpurun_out
for prec in bf16x2 bf16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_h_$prec.csv python bench.py --once --precision $prec > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches_h_$prec.csv > gpurun_out/launch_summary_h_$prec.txt 2>&1; head -12 gpurun_out/launch_summary_h_$prec.txt
done
