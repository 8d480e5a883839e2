#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/tc_errors.log
timeout 1500 python -m pytest tests/test_gpu_tc.py -q -m gpu -x > gpurun_out/pytest_tc.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_tc.log
tail -6 gpurun_out/pytest_tc.log
for prec in bf16x2 bf16; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$prec.csv python bench.py --once --precision $prec > gpurun_out/ncu_list.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$prec.csv > gpurun_out/launch_summary_$prec.txt 2>&1; head -8 gpurun_out/launch_summary_$prec.txt
done
