#!/bin/bash
# Runs the GPU test files one process each (a trapped kernel poisons its CUDA context) with per-file timeouts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rm -f gpurun_out/summary.txt
for f in tests/test_gpu_kernels.py tests/test_gpu_parity.py tests/test_gpu_spinnet.py tests/test_gpu_preproc.py tests/test_reference_modeling.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu -s --timeout 600 > gpurun_out/$n.log 2>&1
  echo "$n exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$n.log
done
