#!/bin/bash
# final profiling trip: launch lists (duration + DRAM bytes) of one forward step and one training step, plus
# --set full captures of the dominant kernels. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
if [ "$1" != "train" ]; then
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_fwd.csv \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_list_fwd.log 2>&1
tail -1 gpurun_out/ncu_list_fwd.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 107 -c 2 -f -o gpurun_out/prof_gemm_r01 \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_ce_partial -c 1 -f -o gpurun_out/prof_loss_r01 \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 24 -c 2 -f -o gpurun_out/prof_attn_r01 \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_full3.log 2>&1
fi
if [ "$1" != "fwd" ]; then
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_list_train.log 2>&1
tail -1 gpurun_out/ncu_list_train.log
fi
ls -la gpurun_out | tail -12
