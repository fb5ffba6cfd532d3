#!/bin/bash
# full GPU check: all parity tests, kernel micro-bench, bench.py (N=1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m "gpu and not slow" -q -x 2>&1 | tail -15 > gpurun_out/t_all.log
cat gpurun_out/t_all.log
if [ "$1" != "notslow" ]; then
timeout 900 python -m pytest tests -m "gpu and slow" -q -s 2>&1 | tail -8 > gpurun_out/t_slow.log
cat gpurun_out/t_slow.log
fi
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
grep -E '"cg": 2, |kd_ce|error' gpurun_out/bench_kernels.log | grep -v '"bn": 128'
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
