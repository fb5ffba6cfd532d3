#!/bin/bash
# round 2, trip R (8 GPUs): training step with the communication stream at the highest priority; NCCL_MAX_CTAS 8 vs 16
mkdir -p gpurun_out
for c in 8 16; do
  NCCL_MAX_CTAS=$c timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2951$((c % 10)) bench.py --gpus 8 --workload train --steps 5 --warmup 3 --no-library-baseline --no-cpu-baseline > gpurun_out/r2r_train_n8_prio_ctas$c.json 2> gpurun_out/r2r_$c.err
  tail -1 gpurun_out/r2r_$c.err | cut -c1-300
done
python - <<'P'
import json
for c in (8,16):
    try:
        d=json.loads(open('gpurun_out/r2r_train_n8_prio_ctas%d.json'%c).read().strip().splitlines()[-1])
        a=d['allreduce_ms']; print('ctas',c, round(d['value'],1), round(d['ms_per_step'],2), 'span',round(a['span'],1),'exposed',round(a['exposed'],2), d['check']['grad_parity']['rel_l2'], d['clocks'])
    except Exception as e: print(c,'ERR',e)
P
