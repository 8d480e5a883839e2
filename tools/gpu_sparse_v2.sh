#!/bin/bash
# sparse (cfg5) visit for the second-generation kernels: stage-by-stage parity, A/B bench, launch list, one full ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_fused.py -q -x  > gpurun_out/pytest_sparse_v2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sparse_v2.log
tail -25 gpurun_out/pytest_sparse_v2.log
for cfgopt in ${BENCH_OPTS:-"sparse_v2_tc=1" "sparse_v2_tc=0" "sparse_v2=0"}; do
  timeout 400 python bench.py --workload cfg5 --batch 64 --steps 2 --warmup 1 --opt $cfgopt > gpurun_out/bench_cfg5_$cfgopt.json 2> gpurun_out/bench_cfg5_$cfgopt.err
  echo "== $cfgopt"; cut -c1-260 gpurun_out/bench_cfg5_$cfgopt.json; tail -3 gpurun_out/bench_cfg5_$cfgopt.err | grep -v Warn | grep -v sparse_csr
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --once > gpurun_out/ncu_list_cfg5.log 2>&1
python tools/launch_summary.py gpurun_out/launches_cfg5.csv > gpurun_out/launch_summary_cfg5.txt 2>&1
cat gpurun_out/launch_summary_cfg5.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'v2_k|gather_contract_k' -s 124 -c 9 -f -o gpurun_out/sp_v2 python bench.py --workload cfg5 --once > gpurun_out/ncu_sp_v2.log 2>&1
ncu -i gpurun_out/sp_v2.ncu-rep --page raw --csv > gpurun_out/sp_v2.raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/sp_v2.raw.csv > gpurun_out/sp_v2_summary.txt 2>&1
cat gpurun_out/sp_v2_summary.txt
