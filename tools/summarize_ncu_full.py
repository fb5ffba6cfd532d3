"""Summarise `ncu --set full` captures (gpurun_out/prof_*_r01.ncu-rep) into the handful of metrics the docs quote."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__cluster_size", "smsp__inst_executed.sum", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__maximum_warps_per_active_cycle_pct"]

for label, path in [a.split("=", 1) for a in sys.argv[1:]]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for n, r in enumerate(rows[2:]):
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## {label} capture {n}: {d.get('Kernel Name', '')[:110]}")
        for k in hdr:
            if any(k.endswith(w) or k == w for w in WANT) and d[k] not in ("", "no data"):
                print(f"  {k} = {d[k]} {u[k]}")
