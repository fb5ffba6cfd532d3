#!/bin/bash
# round 2, trip I: GEMM tail split -- correctness (bounded), microbench A/B, conv0_bwd after the F16 template, step A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x -m gpu -k "tail_split or resid or plain_f32" 2>&1 | tail -5
timeout 300 python tools/bench_gemm_tail.py 2>&1 | tee gpurun_out/r2i_gemm_tail.txt
timeout 120 python tools/bench_conv0_bwd.py 2>&1 | head -3
for v in 1 0 1 0; do B2S_GEMM_TAIL_SPLIT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-train-block >> gpurun_out/r2i_fwd_tail$v.jsonl 2>> gpurun_out/r2i.err; done
for v in 1 0; do B2S_GEMM_TAIL_SPLIT=$v timeout 600 python bench.py --workload train --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline >> gpurun_out/r2i_train_tail$v.jsonl 2>> gpurun_out/r2i.err; done
python - <<'P'
import json
for w in ('fwd','train'):
    for v in (1,0):
        for l in open('gpurun_out/r2i_%s_tail%d.jsonl'%(w,v)):
            d=json.loads(l); print(w,'tail',v, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2))
P
tail -5 gpurun_out/r2i.err
