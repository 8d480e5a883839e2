#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sparse_fused.py tests/test_gpu_train.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/pytest_parity.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_parity.log
grep -n "FAILED\|passed\|failed\|pytest exit" gpurun_out/pytest_parity.log | head -30
for wl in cfg2-edge cfg2-node cfg1; do
timeout 600 python bench.py --workload $wl --no-parity > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; tail -2 gpurun_out/bench_$wl.err | grep -v Warn; python -c "
import json; d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], {k: round(v) if isinstance(v, float) else v for k, v in (d.get('whole_step') or {}).items() if k != 'what'})"
done
