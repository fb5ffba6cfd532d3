#!/bin/bash
# regulariser round trip: new tests, then the training bench deterministic / dropout+SpecAugment / + LayerDrop
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x -k "validate or regularis or train_mode or drop_mask" 2>&1 | tail -8
for r in none dropout all; do
  timeout 600 python bench.py --workload train --steps 4 --warmup 3 --regularize $r > gpurun_out/bench_train_reg_$r.json 2> gpurun_out/bench_train_reg_$r.err
  tail -2 gpurun_out/bench_train_reg_$r.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_train_reg_$r.json"))
print("$r", round(d["value"],1), "utt/s", round(d["ms_per_step"],2), "ms/step e2e", round(d["e2e"]["value"],1), "gemm frac", round(d["roofline"]["frac"],3), "loss", d["check"])
PY
done
