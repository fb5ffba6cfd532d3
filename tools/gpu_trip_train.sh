#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_list_train.log 2>&1
tail -2 gpurun_out/ncu_list_train.log
python tools/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launch_summary_train.txt
cat gpurun_out/launch_summary_train.txt
