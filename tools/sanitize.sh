#!/bin/bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / TMA / 2-CTA kernels on tiny shapes:
# memcheck (out-of-bounds / misaligned global + shared accesses) and racecheck (shared-memory hazards) on
#   * __graft_entry__.smoke(): forward losses, a training micro-batch (LLM + encoder backward), the regularised
#     forward, prefill + KV-cache decode -- every kernel family of the library on tiny HuBERT / Llama shapes;
#   * the GEMM unit shapes (every tile configuration, edge tiles, the TMA-store epilogue) and attention fwd / bwd.
# Summaries go to gpurun_out/ (copy the ones to be judged into profiles/).
mkdir -p gpurun_out
SAN="compute-sanitizer --print-limit 20 --error-exitcode 9"
run() {  # name, tool, command...
  name=$1; tool=$2; shift 2
  timeout 1500 $SAN --tool $tool "$@" > gpurun_out/sanitize_${name}_${tool}.log 2>&1
  echo "== $name / $tool: exit $? ; $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${name}_${tool}.log) summary line(s)"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke OK|passed|failed|Error|Hazard" gpurun_out/sanitize_${name}_${tool}.log | sort | uniq -c | sort -rn | head -12
}
KSEL='test_gemm_plain_f32 and (130-264-200 or 128-256-64) or test_gemm_f16_needs or test_gemm_swiglu or test_gemm_rope or test_gemm_grouped or test_gemm_resid'
ASEL='test_attention and (lens1 or lens4 or lens7) or test_attention_backward and (lens1 or lens4)'
for tool in memcheck racecheck; do
  run smoke $tool python -c "import __graft_entry__ as g; g.smoke()"
  run gemm $tool python -m pytest tests/test_gemm_gpu.py -q -x -k "$KSEL"
  run attn $tool python -m pytest tests/test_ops_gpu.py -q -x -k "$ASEL"
done
