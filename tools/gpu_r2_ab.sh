#!/bin/bash
# cfg5 (configs[4]: sparse kNN graph, batch-sharded on 8 B200) through torchrun
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --workload cfg5 --steps 5 --warmup 2 > gpurun_out/bench_cfg5_n$N.json 2> gpurun_out/bench_cfg5_n$N.err; tail -2 gpurun_out/bench_cfg5_n$N.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5_n$N.json').read().strip().splitlines()[-1]); print('cfg5 N=$N', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), d['config']['global_batch'], d['config']['per_gpu_batch'], d['clocks'])"
