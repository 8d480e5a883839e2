#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
for w in cfg1 cfg2-node cfg2-edge; do python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json; tail -2 gpurun_out/bench_$w.err; done
