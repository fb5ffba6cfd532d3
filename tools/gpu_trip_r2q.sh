#!/bin/bash
# round 2, trip Q: whole GPU suite (incl. the full-size tests), smoke, then the three bench workloads as the driver runs them
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_bench_n1_default_final.json 2> gpurun_out/r2q_default.err; tail -2 gpurun_out/r2q_default.err
timeout 600 python bench.py --workload train --no-library-baseline > gpurun_out/r02_bench_n1_train_final.json 2> gpurun_out/r2q_train.err
timeout 600 python bench.py --workload infer --no-library-baseline > gpurun_out/r02_bench_n1_infer_final.json 2> gpurun_out/r2q_infer.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r2q_ref.err
python - <<'P'
import json
for n in ('default','train','infer'):
    try:
        d=json.loads(open('gpurun_out/r02_bench_n1_%s_final.json'%n).read().strip().splitlines()[-1])
        print(n, round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['clocks'], round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])
        if 'train' in d: t=d['train']; print('   train block', round(t['value'],1), round(t['ms_per_step'],2), t.get('grad_scaler'))
        if 'library_baseline' in d: print('   library', d['library_baseline'].get('value'))
        if 'roofline_loss' in d: print('   loss', round(d['roofline_loss']['frac'],3), d['roofline_loss'].get('traffic'))
        if 'cpu_baseline' in d: print('   cpu', d['cpu_baseline'].get('value'))
    except Exception as e: print(n, 'ERR', e)
print(open('gpurun_out/r02_bench_reference_arm.json').read()[:400])
P
