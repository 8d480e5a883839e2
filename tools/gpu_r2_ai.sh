#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) over the persistent kernels: golden fixtures of every gating mode
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -k "persistent_path_golden" -q -m gpu -p no:cacheprovider > gpurun_out/racecheck_persist.log 2>&1; echo "rc=$?"
grep "passed\|failed\|RACECHECK SUMMARY\|hazard" gpurun_out/racecheck_persist.log | sort | uniq -c | sort -rn | head -12
grep -B2 -A12 "Race reported\|WARN\|ERROR" gpurun_out/racecheck_persist.log | head -60
