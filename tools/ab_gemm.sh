#!/bin/bash
# A/B of the GEMM kernel: current tree vs the tree at _ab_old (same box, back to back, twice)
mkdir -p gpurun_out
for rep in 1 2; do
  (cd _ab_old && timeout 300 python tools/bench_kernels.py 2>&1 | grep '"cg": 2' | grep '"bn": 256' | sed 's/^/OLD /') > gpurun_out/ab_old_$rep.log
  timeout 300 python tools/bench_kernels.py 2>&1 | grep '"cg": 2' | grep '"bn": 256' | sed 's/^/NEW /' > gpurun_out/ab_new_$rep.log
done
python - <<'PY'
import json,glob
res={}
for f in sorted(glob.glob("gpurun_out/ab_*_*.log")):
    for l in open(f):
        tag,js=l.split(" ",1)
        d=json.loads(js)
        res.setdefault(d["label"],{}).setdefault(tag,[]).append(d["tflops"])
for k,v in res.items():
    print(k, {t:x for t,x in v.items()})
PY
