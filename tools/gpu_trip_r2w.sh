#!/bin/bash
# round 2, trip W: gelu_bwd with the fused bias gradient, vectorised AdamW -- tests, training step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py tests/test_decode_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --workload train --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline >> gpurun_out/r2w_train.jsonl 2>> gpurun_out/r2w.err; done
python - <<'P'
import json
for l in open('gpurun_out/r2w_train.jsonl'):
    d=json.loads(l); print('train', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2), d['gpu_launches'])
P
tail -3 gpurun_out/r2w.err
