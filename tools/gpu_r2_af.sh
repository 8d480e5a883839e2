#!/bin/bash
# ncu --set full of the final node-gate kernels and of the fused reverse step with node gates
mkdir -p gpurun_out
cap() {
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name python bench.py "$@" > gpurun_out/ncu_full_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  python tools/ncu_pick.py gpurun_out/$name.raw.csv > gpurun_out/$name.pick.txt 2>&1
  python tools/ncu_source_top.py gpurun_out/$name.source.csv 30 > gpurun_out/$name.top.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep gpurun_out/$name.source.csv
  echo "== $name"; head -4 gpurun_out/$name.pick.txt; grep "issue_active\|warps_active\|registers_per_thread" gpurun_out/$name.pick.txt
}
cap r02_ncu_node_gate_fwd_final node_gate_fwd_kernel 0 --cfg3-spatial node --once
cap r02_ncu_node_gate_bwd_final node_gate_bwd_kernel 0 --cfg3-spatial node --once
cap r02_ncu_bwd_fused_node bwd_fused_kernel 5 --cfg3-spatial node --once
