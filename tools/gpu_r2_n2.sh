#!/bin/bash
# 2 GPUs: data-parallel arithmetic on hardware + bench N=2 through the product path (grad_check)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dist2.py tests/test_gpu_dist_native.py -q -m gpu > gpurun_out/pytest_dist2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_dist2.log
tail -8 gpurun_out/pytest_dist2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 2 --native-allreduce 1 --also "" --no-parity > gpurun_out/bench_n2_native.json 2> gpurun_out/bench_n2_native.err; tail -3 gpurun_out/bench_n2_native.err; cat gpurun_out/bench_n2_native.json | cut -c1-300
