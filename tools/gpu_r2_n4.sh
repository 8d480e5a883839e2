#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err | cut -c1-300; python -c "
import json; d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1]); print('N=$N', round(d['value']), 'e2e', round(d['e2e']['value']), d['modes'].keys(), {k: round(v['value']) for k,v in d['modes'].items()}, d['grad_check'], d['clocks'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 3 --warmup 3 --impl reference > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cat gpurun_out/bench_ref_n$N.json | cut -c1-200
