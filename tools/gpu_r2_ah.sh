#!/bin/bash
# compute-sanitizer memcheck over the new kernels on small cases (persistent kernels in every gating mode and layout, node gates and
# input gradients on the tensor-core path, the renumbered sparse path)
mkdir -p gpurun_out
run() { timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest "$@" -q -m gpu -x -p no:cacheprovider > gpurun_out/sanitizer_$N.log 2>&1; echo "rc=$? $N"; grep -c "ERROR SUMMARY: 0 errors" gpurun_out/sanitizer_$N.log; grep "passed\|failed\|ERROR SUMMARY" gpurun_out/sanitizer_$N.log | tail -3; }
N=persist run tests/test_gpu_parity.py -k "persistent_path"
N=tcnode run tests/test_gpu_tc.py -k "node_gated_cell_matches_fp32_path and 256-32-3-5-8-1 or input_gradients and 256-32-3-4-6-1"
N=reorder run tests/test_gpu_sparse_fused.py -k "library_reorder_vs_oracle"
