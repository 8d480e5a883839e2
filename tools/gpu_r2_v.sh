#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -m gpu -x -k "node_gated or auto_precision or reduced_cfg3" > gpurun_out/pytest_tcnode.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tcnode.log
grep -n "FAILED\|passed\|failed\|pytest exit\|Error\|error" gpurun_out/pytest_tcnode.log | head -20
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_node.csv python bench.py --cfg3-spatial node --once > gpurun_out/ncu_list_node.log 2>&1
python tools/launch_summary.py gpurun_out/launches_node.csv > gpurun_out/launch_summary_node.txt 2>&1; head -12 gpurun_out/launch_summary_node.txt
timeout 1500 python bench.py --cfg3-spatial node --no-secondary --steps 3 --warmup 1 > gpurun_out/bench_cfg3_node.json 2> gpurun_out/bench_cfg3_node.err; tail -3 gpurun_out/bench_cfg3_node.err | grep -v Warn
python -c "
import json; d=json.loads(open('gpurun_out/bench_cfg3_node.json').read().strip().splitlines()[-1]); print('cfg3+node', round(d['value']), round(d['e2e']['value']), d['gpu_launches'], {k: round(v['value']) for k,v in d['modes'].items()}, d['parity']['max_rel_H'], d['parity']['max_rel_grad_T16'])"
