#!/bin/bash
# same-box A/B of an environment toggle: tools/ab_env.sh VAR [extra bench args...]; runs VAR=0 / VAR=1 twice each
VAR=$1; shift
mkdir -p gpurun_out
for i in 1 2; do
  for v in 0 1; do
    env $VAR=$v timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
print("$VAR=$v", "$*", round(d["value"], 1), "utt/s", round(d["ms_per_step"], 2), "ms/step; gemm ms",
      round(d["roofline"]["gemm_ms_per_step"], 2), "sm_mhz", d["clocks"]["sm_mhz"])
PY
  done
done
