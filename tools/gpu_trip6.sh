#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 107 -c 2 -o gpurun_out/prof_gemm_r01 \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:kd_ce_partial -c 1 -o gpurun_out/prof_loss_r01 \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
ls -la gpurun_out
