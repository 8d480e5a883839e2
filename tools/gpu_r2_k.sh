#!/bin/bash
mkdir -p gpurun_out
for o in hilbert random host; do
timeout 900 python bench.py --workload cfg5 --steps 5 --warmup 2 --no-cpu-baseline --cfg5-order $o > gpurun_out/bench_cfg5_$o.json 2> gpurun_out/bench_cfg5_$o.err; tail -1 gpurun_out/bench_cfg5_$o.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5_$o.json').read().strip().splitlines()[-1]); print('$o', round(d['value'],1), round(d['roofline']['frac'],3), d['config']['graph_build_s'])"
done
