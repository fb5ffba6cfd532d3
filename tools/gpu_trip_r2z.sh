#!/bin/bash
# round 2, trip Z: shared prompt prefix in the forward-only paths -- parity, then same-box A/B of the forward step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_path_gpu.py tests/test_configs_gpu.py tests/test_decode_gpu.py tests/test_ops_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_fullsize_gpu.py -q -x -m gpu -k "packed or prefill" -s 2>&1 | grep -E "passed|failed|packed B|decode step|Error|assert" | head -14
for i in 1 2; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-train-block >> gpurun_out/r2z_fwd.jsonl 2>> gpurun_out/r2z.err; done
timeout 600 python bench.py --workload infer --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline >> gpurun_out/r2z_infer.jsonl 2>> gpurun_out/r2z.err
python - <<'P'
import json
for f in ('fwd','infer'):
    for l in open('gpurun_out/r2z_%s.jsonl'%f):
        d=json.loads(l); print(f, round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2), d['config'].get('shared_prefix_rows'))
P
tail -3 gpurun_out/r2z.err
