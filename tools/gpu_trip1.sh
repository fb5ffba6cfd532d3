#!/bin/bash
# first GPU trip: kernel parity per group (separate processes so one trap cannot hide the rest) + micro-bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -q 2>&1 | tail -40 > gpurun_out/t_ops.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "cg1" 2>&1 | tail -40 > gpurun_out/t_gemm_cg1.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "not cg1 and not cg2" 2>&1 | tail -40 > gpurun_out/t_gemm_conv.log
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "cg2" 2>&1 | tail -40 > gpurun_out/t_gemm_cg2.log
timeout 400 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
tail -5 gpurun_out/t_ops.log gpurun_out/t_gemm_cg1.log gpurun_out/t_gemm_conv.log gpurun_out/t_gemm_cg2.log
tail -50 gpurun_out/bench_kernels.log
