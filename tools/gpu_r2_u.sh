#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_node.csv python bench.py --cfg3-spatial node --once > gpurun_out/ncu_list_node.log 2>&1
python tools/launch_summary.py gpurun_out/launches_node.csv > gpurun_out/launch_summary_node.txt 2>&1; head -30 gpurun_out/launch_summary_node.txt
