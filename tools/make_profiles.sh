#!/bin/sh
# Round-end evidence, run on the GPU box via gpurun (one GPU):
#   1. ncu launch list (gpu__time_duration) of a short bench run        -> gpurun_out/launches_bench.csv
#      (tools/launch_agg.py --step 4 cuts the timed step out of it)
#   2. ncu --set full of the top kernels (one launch each)              -> gpurun_out/top_*.ncu-rep
#   3. the bench line of the same build (not under ncu)                 -> gpurun_out/bench_r1.json
# Summaries are produced on the CPU box with tools/summarise_profiles.py -> profiles/
# PARTS selects what to capture (default: everything): a = sweep/covariance/score kernels, b = diag block, c = dgemm
set -x
PARTS=${PARTS:-abc}
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
case $PARTS in *a*)
ncu --set full --clock-control none --import-source on \
    -k regex:"tc_filter_kernel|pair_sweep_kernel|cov_rows_kernel|fn_kernel|pack_planes_kernel|build_lists_kernel|encode_simplex" \
    -c 8 -o gpurun_out/top_a python tools/dev_gpu.py 500x200000 > gpurun_out/ncu_top_a.log 2>&1 ;; esac
case $PARTS in *b*)
ncu --set full --clock-control none --import-source on -k regex:"diag_block_kernel" -c 1 \
    -o gpurun_out/top_b python tools/dev_gpu.py 500x200000 > gpurun_out/ncu_top_b.log 2>&1 ;; esac
# the last dgemm launches of one run: trailing NT updates, the trtri levels, the lauum GEMM
case $PARTS in *c*)
ncu --set full --clock-control none --import-source on -k regex:"dgemm_kernel" -s 186 -c 9 \
    -o gpurun_out/top_c python tools/dev_gpu.py 500x200000 > gpurun_out/ncu_top_c.log 2>&1 ;; esac
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
ls -la gpurun_out
