#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train.py tests/test_gpu_dropin.py -q -m gpu -x > gpurun_out/pytest_parity.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_parity.log
tail -15 gpurun_out/pytest_parity.log
timeout 600 python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; tail -2 gpurun_out/bench_cfg1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg1.json').read().strip().splitlines()[-1]); print('cfg1', round(d['value']), round(d['e2e']['value']), d.get('whole_step'), d['gpu_launches'], d['cpu_baseline'] and round(d['cpu_baseline']['value']))"
timeout 600 python bench.py --workload cfg2-node --no-parity > gpurun_out/bench_cfg2node.json 2> gpurun_out/bench_cfg2node.err; tail -2 gpurun_out/bench_cfg2node.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg2node.json').read().strip().splitlines()[-1]); print('cfg2-node', round(d['value']), round(d['e2e']['value']), d.get('whole_step'), d['gpu_launches'], d['config'])"
