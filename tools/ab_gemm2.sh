#!/bin/bash
# A/B: the library as built (B) vs the 4-epilogue-warp GEMM source in _ab_old/gemm_sm100.cu (A), same box
mkdir -p gpurun_out
echo "== B (new)"; timeout 300 python tools/bench_gemm_enc.py 2>&1 | tee gpurun_out/gemm_B.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x 2>&1 | tail -3
cp _ab_old/gemm_sm100.cu llm_speech_summarization_b200/csrc/gemm_sm100.cu
python -m llm_speech_summarization_b200.build --force > /dev/null 2>&1
echo "== A (old)"; timeout 300 python tools/bench_gemm_enc.py 2>&1 | tee gpurun_out/gemm_A.log
