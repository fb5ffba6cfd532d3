"""Attention forward + backward micro-benchmark only (GPU box)."""
import sys
sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from tools.bench_bwd import bench_attn_bwd  # noqa: E402
from tools.bench_kernels import bench_attn  # noqa: E402

bench_attn("hubert_b32", [499] * 32, 16, 16, 64, False)
bench_attn("llama_b32_student+teacher", [200] * 32 + [117] * 32, 24, 8, 128, True)
bench_attn("minichat_b8_L400", [400] * 8, 24, 24, 128, True)
bench_attn_bwd("hubert 32x499 H16 D64", [499] * 32, 16, 16, 64, False)
bench_attn_bwd("llama 32x200 H24/8 D128 causal", [200] * 32, 24, 8, 128, True)
