#!/bin/bash
# round 2, visit I: whole-step graph test, drop-in tests, small-config benches, full bench x2 with CPU baseline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_dropin.py -q -m gpu > gpurun_out/pytest_train.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_train.log
tail -12 gpurun_out/pytest_train.log
for w in cfg1 cfg2-node cfg2-edge; do
timeout 600 python bench.py --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value']), round(d['e2e']['value']), d.get('whole_step'), d['cpu_baseline'] and round(d['cpu_baseline']['value']))"
done
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -2 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-400
