#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over the kernels added late in round 2: GEMM tail split (flag hand-over), the
# dgrad TMA epilogue (MODE 4), eight epilogue warps (MODE 5), the pipelined attention forward, layernorm_bwd_ex with the
# fused column sum / L2 prefetch, conv0_bwd<F16>, gelu_bwd with the fused bias gradient, the vectorised AdamW + GradScaler.
mkdir -p gpurun_out
SAN="compute-sanitizer --print-limit 20 --error-exitcode 9"
run() {
  name=$1; tool=$2; shift 2
  timeout 200 $SAN --tool $tool "$@" > gpurun_out/sanitize2_${name}_${tool}.log 2>&1
  echo "== $name / $tool: exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|Hazard" gpurun_out/sanitize2_${name}_${tool}.log | sort | uniq -c | sort -rn | head -8
}
GSEL='test_gemm_eight_epilogue_warps or (test_gemm_tail_split_matches_unsplit and dgrad6400_k1024) or (test_gemm_dgrad_mn_major_w and bn256cg2) or test_gemm_dgrad_wgrad_f16'
OSEL='(test_attention_pipelined and (lens1 or lens3 or lens6)) or test_attention_pipelined_lse or (test_layernorm_bwd_ex_vs_autograd and f16) or test_conv0_bwd_vs_autograd or test_gelu_bwd_with_fused or test_grad_scaler_device_state'
for tool in memcheck racecheck; do
  run gemm $tool python -m pytest tests/test_gemm_gpu.py -q -x -k "$GSEL"
  run ops $tool python -m pytest tests/test_ops_gpu.py -q -x -k "$OSEL"
done
(for f in gpurun_out/sanitize2_*.log; do echo "== $f (tail)"; grep -E "COMPUTE-SANITIZER|passed|failed|SUMMARY" $f | tail -4; done) > gpurun_out/r02_compute_sanitizer_late.txt
cat gpurun_out/r02_compute_sanitizer_late.txt
