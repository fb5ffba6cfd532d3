#!/bin/bash
# round 2, trip J: block-per-row layernorm_bwd_ex with the fused dh column sum -- op tests, training parity, train bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -x -m gpu -k gradients -s 2>&1 | grep -v "^$" | tail -22
for i in 1 2; do timeout 600 python bench.py --workload train --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline >> gpurun_out/r2j_train.jsonl 2>> gpurun_out/r2j.err; done
python - <<'P'
import json
for l in open('gpurun_out/r2j_train.jsonl'):
    d=json.loads(l); print('train', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2))
P
tail -3 gpurun_out/r2j.err
