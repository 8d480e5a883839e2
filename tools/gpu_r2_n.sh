#!/bin/bash
# library-owned node renumbering: tests + cfg5 in the generator's random order with the renumbering on / off, and the Hilbert order
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sparse_fused.py tests/test_gpu_parity.py -q -m gpu -x > gpurun_out/pytest_sparse.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sparse.log
tail -8 gpurun_out/pytest_sparse.log
for cfg in "random auto" "random off" "hilbert auto"; do
  set -- $cfg
  timeout 600 python bench.py --workload cfg5 --cfg5-order $1 --cfg5-reorder $2 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg5_$1_$2.json 2> gpurun_out/bench_cfg5_$1_$2.err
  tail -2 gpurun_out/bench_cfg5_$1_$2.err
  python -c "
import json,sys; d=json.loads(open('gpurun_out/bench_cfg5_$1_$2.json').read().strip().splitlines()[-1]); print('cfg5 $1 $2', round(d['value'],1), round(d['e2e']['value'],1), d['config']['library_reorder'])"
done
