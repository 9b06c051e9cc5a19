#!/bin/sh
# Round-end evidence, run on the GPU box via gpurun (one GPU):
#   1. ncu launch list (gpu__time_duration) of exactly one bench step   -> gpurun_out/launches_bench.csv
#   2. ncu --set full of the top kernels (one launch each)              -> gpurun_out/top_*.ncu-rep
#   3. the bench line of the same build (not under ncu)                 -> gpurun_out/bench_r1.json
# Summaries are produced on the CPU box with tools/summarise_profiles.py -> profiles/
set -x
# kernels before the timed step: synth (1) + 3 warm-up steps x 306 launches + 4 L2-flush fills
SKIP=${SKIP:-923}
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c 306 --csv \
    --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"pair_sweep_kernel|cov_rows_kernel|fn_kernel|pack_planes_kernel|build_lists_kernel" \
    -c 6 -o gpurun_out/top_a python tools/dev_gpu.py 500x200000 > gpurun_out/ncu_top_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"diag_block_kernel" -c 1 \
    -o gpurun_out/top_b python tools/dev_gpu.py 500x200000 > gpurun_out/ncu_top_b.log 2>&1
# the last dgemm launches of one run: trailing NT updates, the trtri levels, the lauum GEMM
ncu --set full --clock-control none --import-source on -k regex:"dgemm_kernel" -s 186 -c 9 \
    -o gpurun_out/top_c python tools/dev_gpu.py 500x200000 > gpurun_out/ncu_top_c.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err
ls -la gpurun_out
