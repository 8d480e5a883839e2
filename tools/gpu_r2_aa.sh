#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -m gpu -k "node_gated or fused_backward or reduced_cfg3 or input_gradients" > gpurun_out/pytest_tcnode.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tcnode.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error\|error" gpurun_out/pytest_tcnode.log | head -20; grep "fused-vs-unfused.*node" gpurun_out/tc_errors.log | cut -c1-400 | head -8
timeout 1500 python bench.py --cfg3-spatial node --no-secondary --steps 3 --warmup 1 > gpurun_out/bench_cfg3_node.json 2> gpurun_out/bench_cfg3_node.err; tail -3 gpurun_out/bench_cfg3_node.err | grep -v Warn
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_node.json').read().strip().splitlines()[-1]); print('cfg3+node', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], {k: round(v['value']) for k,v in d['modes'].items()}, d['parity']['max_rel_H'], d['parity']['max_rel_grad_T16'])"
timeout 900 python bench.py --no-secondary --no-parity --also "" --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1]); print('cfg3 (headline unchanged?)', round(d['value']), d['clocks'])"
