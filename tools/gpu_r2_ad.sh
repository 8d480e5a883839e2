#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_ac.sh
bash tools/gpu_r2_v.sh
