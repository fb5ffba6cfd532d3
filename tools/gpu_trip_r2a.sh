#!/bin/bash
# round 2, trip A: full GPU test suite on the fp16-operand path + precision A/B + forward bench A/B (same box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2a_tests.log; tail -25 gpurun_out/r2a_tests.log
for d in fp16 bf16; do DTYPE=$d timeout 600 python tools/diag_precision.py > gpurun_out/r2a_prec_$d.log 2>&1; done
DTYPE=fp16 ENC_DTYPE=bf16 timeout 600 python tools/diag_precision.py > gpurun_out/r2a_prec_llmfp16_encbf16.log 2>&1
grep -h "operand dtype\|rel err\|cuda" gpurun_out/r2a_prec_*.log | grep -v hidden | head -40
for d in fp16 bf16 fp16 bf16; do python bench.py --steps 10 --warmup 3 --dtype $d --no-cpu-baseline >> gpurun_out/r2a_bench_fwd.jsonl 2>> gpurun_out/r2a_bench.err; done
python - <<'P'
import json
for l in open('gpurun_out/r2a_bench_fwd.jsonl'):
    d=json.loads(l); print(d['dtype'], round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
P
