#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_path_gpu.py -m "gpu and not slow" -q 2>&1 | tail -80 > gpurun_out/t_path.log
timeout 200 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "empty" 2>&1 | tail -10 > gpurun_out/t_ops2.log
timeout 900 python -m pytest tests/test_path_gpu.py -m "gpu and slow" -q 2>&1 | tail -60 > gpurun_out/t_path_full.log
cat gpurun_out/t_path.log gpurun_out/t_ops2.log gpurun_out/t_path_full.log
nproc; free -g | head -2
