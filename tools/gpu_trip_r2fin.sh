#!/bin/bash
# round 2, final trip: whole GPU suite, smoke, default bench line, forward launch list of the final code
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench_n1_default_final.json 2> gpurun_out/r2fin.err; tail -1 gpurun_out/r2fin.err | cut -c1-200
timeout 600 python bench.py --workload infer --no-library-baseline > gpurun_out/r02_bench_n1_infer_final.json 2>> gpurun_out/r2fin.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/r02_bench_n1_default_final.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), d["clocks"], round(d["roofline"]["frac"],3), d["gpu_launches"], d["config"].get("shared_prefix_rows"))
print("loss", round(d["roofline_loss"]["frac"],3), round(d["roofline_loss"]["hot"]["frac"],3))
t=d["train"]; print("train", round(t["value"],1), round(t["ms_per_step"],2), round(t["roofline"]["frac"],3)); print("lib", round(d["library_baseline"]["value"],1), "cpu", round(d["cpu_baseline"]["value"],2))
d=json.loads(open("gpurun_out/r02_bench_n1_infer_final.json").read().strip().splitlines()[-1]); print("infer", round(d["value"],1), round(d["ms_per_step"],2))
P
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 700 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_fwd.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/r02_ncu_list_fwd.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_fwd.csv > gpurun_out/r02_launch_summary.txt; head -8 gpurun_out/r02_launch_summary.txt
