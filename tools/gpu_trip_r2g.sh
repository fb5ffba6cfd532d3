#!/bin/bash
# round 2, trip G: decode megakernel -- parity (tiny + full size), then ms/token A/B against the multi-kernel path;
# and the GEMM tile-order (group_m) sweep on the forward step
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py "tests/test_fullsize_gpu.py::test_config1_prefill_then_four_decode_steps_vs_oracle" -q -x -s 2>&1 | tail -15
for m in 1 0; do B2S_DECODE_MEGA=$m DECODE_BATCHES=1,2,4 timeout 600 python tools/bench_decode.py >> gpurun_out/r2g_decode.jsonl 2>> gpurun_out/r2g.err; done
DECODE_BATCHES=8,32 timeout 600 python tools/bench_decode.py >> gpurun_out/r2g_decode.jsonl 2>> gpurun_out/r2g.err
cat gpurun_out/r2g_decode.jsonl
for g in 8 0 16 32 8 0; do B2S_GEMM_GROUP_M=$g python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-train-block > gpurun_out/r2g_tmp.json 2>> gpurun_out/r2g.err; python - <<P
import json
d=json.loads(open('gpurun_out/r2g_tmp.json').read().strip().splitlines()[-1])
print('group_m=$g', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))
open('gpurun_out/r2g_groupm.txt','a').write('group_m=$g %.1f utt/s %.2f ms/step sm %s MHz gemm frac %.3f\n'%(d['value'],d['ms_per_step'],d['clocks']['sm_mhz'],d['roofline']['frac']))
P
done
