#!/bin/bash
# round 2, trip H: conv0_bwd timing by format / data, per-shape GEMM table of the TRAINING step
mkdir -p gpurun_out
timeout 300 python tools/bench_conv0_bwd.py 2>&1 | tee gpurun_out/r2h_conv0_bwd.txt
timeout 600 python bench.py --workload train --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline --gemm-shapes gpurun_out/r2h_shapes_train.txt > gpurun_out/r2h_train.json 2> gpurun_out/r2h.err
tail -3 gpurun_out/r2h.err; cat gpurun_out/r2h_train.json | cut -c1-600
head -60 gpurun_out/r2h_shapes_train.txt
