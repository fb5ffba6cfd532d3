#!/bin/bash
# round 2, trip O: attention backward single-path masking -- tests, then forward / training step on one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py tests/test_gemm_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-train-block >> gpurun_out/r2o_fwd.jsonl 2>> gpurun_out/r2o.err; done
for i in 1 2; do timeout 600 python bench.py --workload train --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline >> gpurun_out/r2o_train.jsonl 2>> gpurun_out/r2o.err; done
python - <<'P'
import json
for w in ('fwd','train'):
    for l in open('gpurun_out/r2o_%s.jsonl'%w):
        d=json.loads(l); print(w, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2))
P
tail -3 gpurun_out/r2o.err
