#!/bin/bash
# ncu --set full of one intermediate Horner stage (bf16 mode)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hshift -s 9 -c 1 -f -o gpurun_out/hshift python bench.py --once --precision bf16 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/hshift.ncu-rep --page raw --csv > gpurun_out/hshift.raw.csv 2>/dev/null
ncu -i gpurun_out/hshift.ncu-rep --page source --csv > gpurun_out/hshift.source.csv 2>/dev/null
python tools/ncu_source_top.py gpurun_out/hshift.source.csv 40
