#!/bin/bash
# ncu launch lists (gpu__time_duration) of one micro-batch forward + backward of the final code: cfg3 in both operand modes, cfg1, cfg2-edge
mkdir -p gpurun_out
for prec in bf16x2 bf16; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_final_$prec.csv python bench.py --once --precision $prec > gpurun_out/ncu_list_final_$prec.log 2>&1
python tools/launch_summary.py gpurun_out/launches_final_$prec.csv > gpurun_out/launch_summary_final_$prec.txt 2>&1; echo "== $prec"; head -9 gpurun_out/launch_summary_final_$prec.txt
done
for wl in cfg1 cfg2-edge; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final_$wl.csv python bench.py --workload $wl --no-cpu-baseline --no-whole-step --no-parity --opt graph_capture=0 --steps 3 --warmup 1 > gpurun_out/ncu_list_final_$wl.log 2>&1
python tools/launch_summary.py gpurun_out/launches_final_$wl.csv > gpurun_out/launch_summary_final_$wl.txt 2>&1; echo "== $wl"; head -6 gpurun_out/launch_summary_final_$wl.txt
done
rm -f gpurun_out/launches_final_*.csv
