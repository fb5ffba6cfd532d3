#!/bin/bash
# round 2 profiling trip: launch lists (duration + DRAM bytes) of one forward step and one training step of bench.py, plus
# --set full captures of the dominant kernels. Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
if [ "$1" != "train" ]; then
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_fwd.csv \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_ncu_list_fwd.log 2>&1
tail -1 gpurun_out/r02_ncu_list_fwd.log
python tools/summarize_launches.py gpurun_out/r02_launches_fwd.csv > gpurun_out/r02_launch_summary.txt; head -12 gpurun_out/r02_launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 107 -c 2 -f -o gpurun_out/r02_prof_gemm \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kd_ce_partial -c 1 -f -o gpurun_out/r02_prof_loss \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_ncu_full2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 24 -c 2 -f -o gpurun_out/r02_prof_attn \
    python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_ncu_full3.log 2>&1
fi
if [ "$1" != "fwd" ]; then
timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_ncu_list_train.log 2>&1
tail -1 gpurun_out/r02_ncu_list_train.log
python tools/summarize_launches.py gpurun_out/r02_launches_train.csv > gpurun_out/r02_launch_summary_train.txt; head -30 gpurun_out/r02_launch_summary_train.txt
fi
ls -la gpurun_out | tail -12
