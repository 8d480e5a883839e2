#!/bin/bash
# ncu --set full of one launch of each fused sparse kernel (cfg5 shape, one micro-batch)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'spmm32_k|filter_fwd_k|aggregate_k|rowstats_k' -s 8 -c 4 -f -o gpurun_out/sp_fwd python bench.py --workload cfg5 --once > gpurun_out/ncu_sp_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bwd_rows_k|bwd_node_k|dh_k|dpre_k' -s 8 -c 4 -f -o gpurun_out/sp_bwd python bench.py --workload cfg5 --once > gpurun_out/ncu_sp_bwd.log 2>&1
for f in sp_fwd sp_bwd; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
tail -3 gpurun_out/ncu_sp_fwd.log gpurun_out/ncu_sp_bwd.log
