#!/bin/bash
# ncu --set full of the current persistent kernels (cfg1, cfg2-edge) and of the node-gate / dpre kernels of the tensor-core path
mkdir -p gpurun_out
cap() {   # name regex skip workload-args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name python bench.py "$@" > gpurun_out/ncu_full_$name.log 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
  python tools/ncu_pick.py gpurun_out/$name.raw.csv > gpurun_out/$name.pick.txt 2>&1
  python tools/ncu_source_top.py gpurun_out/$name.source.csv 30 > gpurun_out/$name.top.txt 2>&1
  rm -f gpurun_out/$name.ncu-rep gpurun_out/$name.source.csv
  echo "== $name"; head -3 gpurun_out/$name.pick.txt
}
cap r02_ncu_persist_fwd_cfg1 persist_fwd_k 3 --workload cfg1 --no-cpu-baseline --no-whole-step --opt graph_capture=0
cap r02_ncu_persist_bwd_cfg1 persist_bwd_k 3 --workload cfg1 --no-cpu-baseline --no-whole-step --opt graph_capture=0
cap r02_ncu_persist_bwd_cfg2edge persist_bwd_k 3 --workload cfg2-edge --no-cpu-baseline --no-whole-step --opt graph_capture=0
cap r02_ncu_node_gate_fwd node_gate_fwd_kernel 0 --cfg3-spatial node --once
cap r02_ncu_node_gate_bwd node_gate_bwd_kernel 0 --cfg3-spatial node --once
cap r02_ncu_dpre_node dpre_kernel 5 --cfg3-spatial node --once
