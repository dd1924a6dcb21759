#!/bin/bash
# ncu --set full for the HBM-/latency-bound kernels (north_star: "achieved HBM GB/s on the gather/kNN stages").
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'ln_modulate|head_final|posenc|procrustes_kernel|patch_kernel|voxel_feature|pool_kernel|fps_kernel' -c 40 -o gpurun_out/prof_hbm -f \
    python tools/profile_hbm_kernels.py > gpurun_out/ncu_hbm.log 2>&1
echo "hbm rc=$?"
ls -la gpurun_out/prof_hbm.ncu-rep
