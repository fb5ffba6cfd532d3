#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py -q -x -m gpu -k "attention or regular or drop" 2>&1 | tail -2
timeout 200 python tools/bench_kernels.py attn 2>&1 | grep -v "bf16" | tee gpurun_out/r2n_attn.jsonl
