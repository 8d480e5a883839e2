#!/bin/bash
# ncu --set full of the shift GEMM at bench.py's launch shapes (P = 2 and P = 1), for roofline.traffic
mkdir -p gpurun_out
for prec in bf16x2 bf16; do
P=2; [ $prec = bf16 ] && P=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:shift_gemm2 -s 40 -c 1 -f -o gpurun_out/gemm2_$prec python bench.py --once --precision $prec > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/gemm2_$prec.ncu-rep --page raw --csv > gpurun_out/r02_ncu_gemm2_R131072_P$P.raw.csv 2>/dev/null
python tools/ncu_pick.py gpurun_out/r02_ncu_gemm2_R131072_P$P.raw.csv | head -6
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-secondary > gpurun_out/bench_modes.json 2> gpurun_out/bench_modes.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_modes.json').read().strip().splitlines()[-1]); print(d['modes'])"
