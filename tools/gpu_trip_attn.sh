#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -x -k "attention or attn" 2>&1 | tail -6
timeout 300 python tools/bench_kernels.py attn 2>&1 | tee gpurun_out/bench_attn.log
timeout 600 python -m pytest tests/test_path_gpu.py tests/test_train_gpu.py -q -x 2>&1 | tail -5
