#!/bin/bash
# all GPU tests, then the forward and training benches (N=1)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fwd_quick.json 2> gpurun_out/bench_fwd_quick.err
timeout 600 python bench.py --workload train --steps 4 --warmup 3 > gpurun_out/bench_train_quick.json 2> gpurun_out/bench_train_quick.err
python - <<PY
import json
for n in ("fwd","train"):
    try:
        d=json.load(open(f"gpurun_out/bench_{n}_quick.json"))
        print(n, round(d["value"],1), "utt/s", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["value"],1), "gemm frac", round(d["roofline"]["frac"],3), "gemm share", round(d["roofline"].get("gemm_share_of_step",0),3), d["clocks"])
    except Exception as e:
        print(n, "failed", e); print(open(f"gpurun_out/bench_{n}_quick.err").read()[-2000:])
PY
