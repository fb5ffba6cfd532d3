#!/bin/bash
# round 2, trip M: pipelined attention forward (S double-buffered) -- bounded correctness run, then kernel A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -m gpu -k "attention" 2>&1 | tail -8
timeout 200 python tools/bench_kernels.py attn 2>&1 | tee gpurun_out/r2m_attn.jsonl
