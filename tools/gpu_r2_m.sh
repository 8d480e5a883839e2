#!/bin/bash
# ncu --set full + source view of the two persistent small-graph kernels at cfg1
mkdir -p gpurun_out
for k in persist_fwd_k persist_bwd_k; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/$k python bench.py --workload cfg1 --no-cpu-baseline --no-whole-step --opt graph_capture=0 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/$k.ncu-rep --page raw --csv > gpurun_out/$k.raw.csv 2>/dev/null
ncu -i gpurun_out/$k.ncu-rep --page source --csv > gpurun_out/$k.source.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/$k.raw.csv > gpurun_out/$k.pick.txt 2>&1
python tools/ncu_source_top.py gpurun_out/$k.source.csv 45 > gpurun_out/$k.top.txt 2>&1
head -3 gpurun_out/$k.pick.txt
done
