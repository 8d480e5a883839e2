#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "reduced_cfg3" > gpurun_out/pytest_tcnode2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tcnode2.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error\|error" gpurun_out/pytest_tcnode2.log | head -20; grep "tc-vs-fp64" gpurun_out/tc_errors.log | head
timeout 1500 python bench.py --cfg3-spatial node --no-secondary --steps 3 --warmup 1 > gpurun_out/bench_cfg3_node.json 2> gpurun_out/bench_cfg3_node.err; tail -3 gpurun_out/bench_cfg3_node.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_node.json').read().strip().splitlines()[-1]); print('cfg3+node', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], d.get('modes'), d.get('parity'))"
