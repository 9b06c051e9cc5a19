#!/bin/sh
# Round-2 evidence of the final build (one GPU, via gpurun).  Summaries: python tools/summarise_profiles.py r2 (CPU box).
set -x
ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py C > gpurun_out/r2_step_under_ncu.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"cov_tc_kernel|tc_filter_kernel|cell_sweep_kernel|encode_onehot4_kernel|site_hist_kernel|diag_block_kernel2|dgemm_small_kernel" -c 8 \
    -o gpurun_out/r2_top_b python tools/profile_step.py C > gpurun_out/r2_ncu_b.log 2>&1
# the LAST sliced GEMM launches of the step: the h = 64 trtri level and the lauum product
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ozaki_gemm_kernel" -s 81 -c 2 \
    -o gpurun_out/r2_top_oz python tools/profile_step.py C > gpurun_out/r2_ncu_oz.log 2>&1
ls -la gpurun_out | tail -8
