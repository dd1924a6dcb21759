#!/bin/bash
# ncu evidence for profiles/: (1) per-launch durations of one full pair, (2) full-set captures of the hot kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_pair.py > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
FLOW_STEPS=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:fmha -c 2 -o gpurun_out/prof_fmha -f \
    python tools/profile_pair.py > gpurun_out/ncu_fmha.log 2>&1
echo "fmha rc=$?"
FLOW_STEPS=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 5 -o gpurun_out/prof_gemm -f \
    python tools/profile_pair.py > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
