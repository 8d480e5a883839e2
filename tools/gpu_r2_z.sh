#!/bin/bash
# 2 GPUs: smoke, data-parallel tests on hardware, cfg5 at N = 2 (batch sharded)
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 1200 python -m pytest tests/test_gpu_dist2.py tests/test_gpu_dist_native.py -q -m gpu > gpurun_out/pytest_dist2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist2.log
grep -n "passed\|failed\|pytest exit" gpurun_out/pytest_dist2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --workload cfg5 --steps 3 --warmup 1 > gpurun_out/bench_cfg5_n2.json 2> gpurun_out/bench_cfg5_n2.err; tail -2 gpurun_out/bench_cfg5_n2.err | cut -c1-200
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg5_n2.json').read().strip().splitlines()[-1]); print('cfg5 N=2', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'])"
