#!/bin/sh
# Final evidence of a build, run on the GPU box via gpurun (one GPU): GPU test suite, smoke, launch lists of one step (C and D), bench line.
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv python tools/profile_step.py C > gpurun_out/r2_step_under_ncu.log 2>&1
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv \
    --log-file gpurun_out/r2_launches_D.csv python tools/profile_step.py D > gpurun_out/r2_step_D_under_ncu.log 2>&1
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_final_nb.json 2> gpurun_out/bench_final_nb.err; tail -c 600 gpurun_out/bench_final_nb.json
