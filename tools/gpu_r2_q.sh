#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sparse_fused.py tests/test_gpu_parity.py -q -m gpu > gpurun_out/pytest_sparse.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_sparse.log
grep -n "FAILED\|passed\|failed\|pytest exit\|deviates" gpurun_out/pytest_sparse.log | head; cat gpurun_out/cfg5_kink_calibration.json
for wl in cfg1 cfg2-node cfg2-edge; do
timeout 600 python bench.py --workload $wl --no-parity > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], {k: round(v) if isinstance(v, float) else v for k, v in (d.get('whole_step') or {}).items() if k != 'what'})"
done
