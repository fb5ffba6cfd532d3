#!/bin/bash
# round 2, trip P: TMA epilogue for the backward GEMMs (dgrad store, wgrad / split-K reduce-add) -- tests, A/B of the training step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_train_gpu.py tests/test_decode_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -3
for v in 1 0 1 0; do B2S_TMA_EPI=$v timeout 600 python bench.py --workload train --steps 6 --warmup 3 --no-cpu-baseline --no-library-baseline --gemm-shapes gpurun_out/r2p_shapes_tma$v.txt >> gpurun_out/r2p_train_tma$v.jsonl 2>> gpurun_out/r2p.err; done
python - <<'P'
import json
for v in (1,0):
    for l in open('gpurun_out/r2p_train_tma%d.jsonl'%v):
        d=json.loads(l); print('train tma',v, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2))
P
for v in 1 0; do echo "== tma=$v"; grep -E "dgrad|wgrad" gpurun_out/r2p_shapes_tma$v.txt | head -14; done
tail -3 gpurun_out/r2p.err
