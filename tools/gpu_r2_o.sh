#!/bin/bash
# launch lists of cfg5 (one micro-batch forward + backward) in random order with the library renumbering, and in Hilbert order
mkdir -p gpurun_out
for cfg in "random auto" "hilbert auto" "random off"; do
  set -- $cfg
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg5_$1_$2.csv python bench.py --workload cfg5 --cfg5-order $1 --cfg5-reorder $2 --once > gpurun_out/ncu_list_cfg5_$1_$2.log 2>&1
  python tools/launch_summary.py gpurun_out/launches_cfg5_$1_$2.csv > gpurun_out/launch_summary_cfg5_$1_$2.txt 2>&1
  echo "== $1 $2"; head -24 gpurun_out/launch_summary_cfg5_$1_$2.txt
done
