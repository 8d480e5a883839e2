#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench.py default run: $SECONDS s"; tail -3 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'modes', {k: round(v['value']) for k,v in d['modes'].items()})
print('roofline', {k: d['roofline'][k] for k in ('frac','tensor_pipe_frac','step_frac','step_frac_of_mode_equivalent_peak')})
print('parity', {k: d['parity'][k] for k in ('max_rel_H_first16','max_rel_H','max_rel_grad_T16')})
print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'clocks', d['clocks'])
for k,v in (d.get('secondary') or {}).items(): print('secondary', k, {kk: (vv if not isinstance(vv, dict) else '...') for kk,vv in v.items() if kk in ('value','ms_per_step','error','steps')}, v.get('roofline') and v['roofline'].get('frac'), v.get('whole_step') and v['whole_step'].get('graphed_seq_per_s'))
PY
