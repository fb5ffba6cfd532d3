#!/bin/bash
# round 2, trip F: ncu --set full of one FFN1+GELU GEMM (15968 x 4096 x 1024) and one out-projection (15968 x 1024 x 1024)
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on -f"
timeout 600 ncu $COMMON -k regex:gemm_bf16 -s 10 -c 1 -o gpurun_out/r2f_ffn1 python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r2f_ffn1.log 2>&1; tail -2 gpurun_out/r2f_ffn1.log
timeout 600 ncu $COMMON -k regex:gemm_bf16 -s 9 -c 1 -o gpurun_out/r2f_outproj python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r2f_outproj.log 2>&1; tail -2 gpurun_out/r2f_outproj.log
ls -la gpurun_out/*.ncu-rep
