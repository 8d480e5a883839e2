#!/bin/bash
# same-box scaling check: N = 1 and N = 8 back to back (bf16x2 only, no side legs)
mkdir -p gpurun_out
F="--steps 5 --warmup 3 --no-secondary --no-parity --also  --no-cpu-baseline"
timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary --no-parity --also "" --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 5 --warmup 3 --no-secondary --no-parity --also "" --no-cpu-baseline > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
python - <<'PY'
import json
base=None
for N in (1,2,4,8):
    d=json.loads(open(f'gpurun_out/scale_n{N}.json').read().strip().splitlines()[-1])
    if N==1: base=d['value']; be=d['e2e']['value']
    print(N, round(d['value']), 'eff', round(d['value']/(N*base),3), 'e2e', round(d['e2e']['value']), 'eff', round(d['e2e']['value']/(N*be),3), 'gemm', round(d['roofline']['tensor_pipe_frac'],3), d['clocks'], d['config']['microbatch'])
PY
