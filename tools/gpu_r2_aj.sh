#!/bin/bash
mkdir -p gpurun_out
run() { timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest "$@" -q -m gpu -x -p no:cacheprovider > gpurun_out/sanitizer_$N.log 2>&1; echo "rc=$? $N"; grep "passed\|failed\|ERROR SUMMARY" gpurun_out/sanitizer_$N.log | tail -2; }
N=reorder8k run tests/test_gpu_sparse_fused.py -k "reorder_decision or reorder_with_input"
N=tcfusednode run tests/test_gpu_tc.py -k "fused_backward_step and node and 256-4-3-5-2"
