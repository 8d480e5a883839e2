#!/bin/bash
# full validation: whole GPU suite, smoke, both bench arms (default flags, as the driver runs them)
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-300
SECONDS=0; timeout 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench.py default run: $SECONDS s"; tail -3 gpurun_out/bench_n1.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'modes', {k: round(v['value']) for k,v in d['modes'].items()})
print('roofline', {k: d['roofline'][k] for k in ('frac','tensor_pipe_frac','step_frac','step_frac_of_mode_equivalent_peak')})
print('parity', {k: d['parity'][k] for k in ('max_rel_H_first16','max_rel_H','max_rel_grad_T16')})
print('cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'clocks', d['clocks'])
for k,v in (d.get('secondary') or {}).items(): print('secondary', k, {kk: (vv if not isinstance(vv, dict) else {a: round(b) if isinstance(b, float) else b for a, b in vv.items() if a != 'workload'}) for kk,vv in v.items() if kk in ('value','ms_per_step','error','steps','bf16x2','bf16')}, v.get('roofline') and v['roofline'].get('frac'), v.get('whole_step') and v['whole_step'].get('graphed_seq_per_s'))
PY
