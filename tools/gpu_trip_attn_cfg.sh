#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/attn_cfgs.log
for cfg in "64,1" "64,2" "128,1" "128,2"; do
  echo "== B2S_ATTN_CFG=$cfg" | tee -a gpurun_out/attn_cfgs.log
  B2S_ATTN_CFG=$cfg timeout 300 python tools/bench_kernels.py attn 2>&1 | grep -E "tcgen05|rror" | tee -a gpurun_out/attn_cfgs.log
done
B2S_ATTN_CFG=64,2 timeout 300 python -m pytest tests/test_ops_gpu.py -q -x -k "attention" 2>&1 | tail -3
