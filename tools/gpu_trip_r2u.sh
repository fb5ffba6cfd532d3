#!/bin/bash
# round 2, trip U: eight epilogue warps (MODE 5) -- the GEMM suite with the option forced, then same-box A/B of the forward step
mkdir -p gpurun_out
B2S_GEMM_EPI8=2 timeout 400 python -m pytest tests/test_gemm_gpu.py -q -x -m gpu 2>&1 | tail -2
timeout 400 python -m pytest tests/test_gemm_gpu.py tests/test_path_gpu.py -q -x -m "gpu and not slow" 2>&1 | tail -2
for v in 1 0 1 0; do B2S_GEMM_EPI8=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-train-block --gemm-shapes gpurun_out/r2u_shapes_e$v.txt >> gpurun_out/r2u_fwd_e$v.jsonl 2>> gpurun_out/r2u.err; done
python - <<'P'
import json
for v in (1,0):
    for l in open('gpurun_out/r2u_fwd_e%d.jsonl'%v):
        d=json.loads(l); print('epi8',v, round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3), round(d['roofline']['gemm_ms_per_step'],2))
P
for v in 1 0; do echo "== epi8=$v"; sed -n 3,16p gpurun_out/r2u_shapes_e$v.txt; done
tail -3 gpurun_out/r2u.err
