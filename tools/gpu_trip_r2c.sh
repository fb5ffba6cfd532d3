#!/bin/bash
# round 2, trip C: full-size parity tests (gradients / packed batch / decode) + the new default bench line + infer line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fullsize_gpu.py -q -s -x 2>&1 | tail -60 > gpurun_out/r2c_fullsize.log; grep -v "^$" gpurun_out/r2c_fullsize.log | tail -45
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench_default.json 2> gpurun_out/r2c_bench_default.err; tail -c 300 gpurun_out/r2c_bench_default.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2c_bench_default.json').read().strip().splitlines()[-1])
print({k:(v if not isinstance(v,dict) else '...') for k,v in d.items()})
print('roofline',d['roofline']['frac'],'loss',d['roofline_loss']['frac'],d['roofline_loss']['hot']['frac'])
print('lib',d['library_baseline']); print('cpu',d['cpu_baseline'])
t=d['train']; print('train',t['value'],t['ms_per_step'],t['e2e']['value'],t['roofline']['frac'],t['allreduce_ms'],t['grad_scaler'],t['check'])
P
timeout 600 python bench.py --workload infer --steps 10 --warmup 3 > gpurun_out/r2c_bench_infer.json 2> gpurun_out/r2c_bench_infer.err; tail -c 300 gpurun_out/r2c_bench_infer.err; head -c 1500 gpurun_out/r2c_bench_infer.json
