#!/bin/bash
# round 2, trip D (2 GPUs): the bucketed / overlapped gradient all-reduce on the real path, A/B against the monolithic one
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --workload train --steps 8 --warmup 3 > gpurun_out/r2d_train_n2_overlap.json 2> gpurun_out/r2d_train_n2_overlap.err; tail -c 400 gpurun_out/r2d_train_n2_overlap.err
timeout 600 $TR bench.py --gpus 2 --workload train --steps 8 --warmup 3 --no-overlap > gpurun_out/r2d_train_n2_mono.json 2> gpurun_out/r2d_train_n2_mono.err; tail -c 400 gpurun_out/r2d_train_n2_mono.err
timeout 900 $TR bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2d_default_n2.json 2> gpurun_out/r2d_default_n2.err; tail -c 400 gpurun_out/r2d_default_n2.err
python - <<'P'
import json
for f in ('r2d_train_n2_overlap','r2d_train_n2_mono'):
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],2), d['allreduce_ms'], d['check']['grad_parity'], d['grad_scaler'])
    except Exception as e: print(f, 'ERR', e)
try:
    d=json.loads(open('gpurun_out/r2d_default_n2.json').read().strip().splitlines()[-1])
    t=d['train']; print('default n2 fwd', round(d['value'],1), 'train', round(t['value'],1), t['ms_per_step'], t['allreduce_ms'], t['check'])
except Exception as e: print('default ERR', e)
P
