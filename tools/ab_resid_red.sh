#!/bin/bash
# same-box A/B of the in-place residual epilogue: B2S_RESID_RED=0 (load + add + store) vs 1 (red.global.add in L2)
mkdir -p gpurun_out
for i in 1 2; do
  for v in 0 1; do
    B2S_RESID_RED=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2>/dev/null
    python - <<PY
import json
d = json.load(open("gpurun_out/ab_$v.json"))
print("RESID_RED=$v", round(d["value"], 1), "utt/s", round(d["ms_per_step"], 2), "ms/step; gemm ms",
      round(d["roofline"]["gemm_ms_per_step"], 2), "sm_mhz", d["clocks"]["sm_mhz"])
PY
  done
done
