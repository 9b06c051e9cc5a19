#!/bin/sh
# Round-2 evidence, run on the GPU box via gpurun (one GPU).  Summaries are made on the CPU box by tools/summarise_profiles_r2.py.
set -x
# 1. launch list of ONE step (warm-up unprofiled): every launch with its device time, cold-cache and serialised -> compare shares
ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py C > gpurun_out/r2_step_under_ncu.log 2>&1
# 2. ncu --set full of the hot kernels, one or two launches each
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"cov_rows_kernel|tc_filter_kernel|pair_sweep_kernel|slice_cols_kernel|slice_rows_kernel" -c 6 \
    -o gpurun_out/r2_top_a python tools/profile_step.py C > gpurun_out/r2_ncu_a.log 2>&1
# the LAST sliced GEMM launches of the step: the h = 32 / 64 trtri levels and the lauum product
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ozaki_gemm_kernel" -s 44 -c 5 \
    -o gpurun_out/r2_top_oz python tools/profile_step.py C > gpurun_out/r2_ncu_oz.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"diag_block_kernel" -c 1 \
    -o gpurun_out/r2_top_diag python tools/profile_step.py C > gpurun_out/r2_ncu_diag.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"di_kernel|site_chol_kernel|fn_kernel" -c 3 \
    -o gpurun_out/r2_top_di python tools/profile_step.py D > gpurun_out/r2_ncu_di.log 2>&1
ls -la gpurun_out | tail -12
