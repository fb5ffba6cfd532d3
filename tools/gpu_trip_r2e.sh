#!/bin/bash
# round 2, trip E: TMA-store GEMM epilogue -- correctness, then same-box A/B (B2S_TMA_EPI=0/1) with per-shape tables
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_path_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -8
for v in 1 0 1 0; do B2S_TMA_EPI=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-train-block --gemm-shapes gpurun_out/r2e_shapes_tma$v.txt >> gpurun_out/r2e_fwd_tma$v.jsonl 2>> gpurun_out/r2e.err; done
python - <<'P'
import json
for v in (1,0):
    for l in open('gpurun_out/r2e_fwd_tma%d.jsonl'%v):
        d=json.loads(l); print('tma',v, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2))
P
for v in 1 0; do echo "== shapes tma=$v"; head -16 gpurun_out/r2e_shapes_tma$v.txt; done
