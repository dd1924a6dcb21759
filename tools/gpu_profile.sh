#!/bin/bash
# ncu evidence for profiles/: (1) per-launch durations of one full pair, (2) full-set captures of the hot kernels.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
# launches before the timed region: weight packing (~430) + set_condition + 3 warm-up pairs (2746 each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 8700 -c 2800 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fmha_kernel -s 60 -c 2 -o gpurun_out/prof_fmha -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fmha.log 2>&1
echo "fmha rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 200 -c 6 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
echo "gemm rc=$?"
ls -la gpurun_out/*.ncu-rep
