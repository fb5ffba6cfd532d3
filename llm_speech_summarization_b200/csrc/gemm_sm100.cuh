// gemm_sm100.cuh -- argument block of the tcgen05 GEMM (see gemm_sm100.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b2s {

enum EpiMode : int {
  EPI_BF16 = 0,       // out_bf16 = act(acc + bias)
  EPI_RESID_F32 = 1,  // out_f32  = resid_f32 + act(acc + bias)          (out may alias resid)
  EPI_SWIGLU = 2,     // out_bf16[:, j] = silu(acc[:, j]) * acc[:, j + BN/2] per BN-wide tile (gate|up packed)
  EPI_ROPE = 3,       // out_bf16 = rotate-half RoPE over 128-wide heads for cols < rope_cols, else passthrough
  EPI_F32 = 4,        // out_f32  = act(acc + bias)
  EPI_ACCUM_F32 = 5,  // out_f32 += acc + bias   (atomic fp32 adds: split-K partials, gradient accumulation)
};

enum ActMode : int { ACT_NONE = 0, ACT_GELU = 1 };

// C[b, m, g*N + n] = epi( sum_k A[b, m (+tap), k] * W[g*w_group_off + n, k] )
// A is any bf16 tensor addressable as a 3-D TMA view (k contiguous, row stride, batch stride): plain
// row-major activations, the overlapping strided views that turn a strided Conv1d into a GEMM, or the
// "tap" walk used by the grouped positional convolution (row coordinate advances with the k-block and
// out-of-range rows are zero-filled by TMA = the conv's zero padding).
struct GemmArgs {
  // A view
  const void* A;
  int a_dim0;               // extent of the contiguous dim (elements)
  long long a_row_stride;   // elements between consecutive rows  (multiple of 8)
  long long a_batch_stride; // elements between batches           (multiple of 8)
  int a_rows;               // rows per batch addressable in the view (TMA zero-fills beyond)
  // W: row-major [w_rows, w_cols] bf16 (nn.Linear layout: out_features x in_features)
  const void* W;
  int w_rows;
  int w_cols;
  // problem
  int M;           // output rows per batch
  int N;           // output cols per group
  int batches;
  int groups;
  int taps;        // 1 for a plain GEMM
  int k_per_tap;   // reduction elements per tap (K for a plain GEMM)
  int a_pad;       // row coordinate = m + tap - a_pad
  int a_group_off; // A column offset per group
  int w_group_off; // W row offset per group
  // epilogue
  int epi;
  int act;
  const float* bias;        // [groups*N] fp32 or null
  void* out;
  long long ldo;            // output leading dim (elements)
  long long out_batch_rows; // output row = b*out_batch_rows + m
  const float* resid;       // EPI_RESID_F32: fp32 [rows, ldo]
  int resid_bcast;          // 1: resid is [M, ldo], shared by every batch (row index = m, not b*out_batch_rows + m)
  void* out2;               // optional bf16 [rows, ld2]: (acc + bias) BEFORE the activation (EPI_BF16) / the raw
  long long ld2;            //   gate|up columns (EPI_SWIGLU) -- what the backward pass needs
  const float* rope_cs;     // EPI_ROPE: [npos, 128] fp32 = cos[0:64] | sin[0:64]
  const int* positions;     // EPI_ROPE: [rows] position of each row inside its own sequence
  int rope_cols;
  // tuning (0 = auto)
  int block_n;    // 64 / 128 / 256
  int cta_group;  // 1 / 2
  // ---- backward GEMMs: MN-major operands (no transposed copies), reduction over batches, split-K
  //   b_mn: W is [w_rows = K, w_cols = N-extent] row-major (dgrad: dX = dY . W with W in nn.Linear layout)
  //   a_mn: A is [a_rows = K, a_dim0 = M-extent] row-major; the reduction additionally runs over k_batches
  //         batches of both operands (wgrad: dW[m, n] = sum_{b, t} dY[b, t, m] * X[b, t, n]); k_per_tap = rows per batch
  int a_mn;
  int b_mn;
  int k_batches;            // 0/1 = none
  long long w_row_stride;   // MN-major W: elements between rows (0 = w_cols) -- overlapping conv windows allowed
  long long w_batch_stride; // MN-major W: elements between k-batches
  int k_splits;             // 0 = auto (EPI_ACCUM_F32 only), 1 = none
  int b_tap_atoms;          // grouped-conv wgrad: N atom j reads W columns [g*w_group_off, +64) at rows k + j - a_pad
  int out_group_rows;       // output row offset per group   (default 0)
  int out_group_cols;       // output column offset per group (default N when out_group_rows == 0)
  // ---- train-mode dropout fused into the epilogue (rng.cuh): applied to act(acc + bias), i.e. AFTER the activation and
  // BEFORE the residual add -- the order of nn.Dropout in HubertFeedForward / HubertEncoderLayerStableLayerNorm.
  // Element index = output_row * ldo + output_col. drop_thresh == 0: off. out2 keeps the un-dropped pre-activation.
  unsigned drop_k1, drop_k2, drop_thresh;
  float drop_inv_keep;
  // ---- 16-bit storage formats (B2S_FMT_BF16 = 0 / B2S_FMT_F16 = 1): A, W and the 16-bit outputs (out, out2).
  // A and W must agree (a mixed bf16 x fp16 tcgen05.mma traps on B200); the output format is free.
  int a_fmt, w_fmt, out_fmt;
};

int gemm_bf16_launch(const GemmArgs& a, cudaStream_t stream);
// the common case: operands and 16-bit outputs all in `fmt`
inline int gemm_launch_fmt(GemmArgs g, int fmt, cudaStream_t stream) {
  g.a_fmt = g.w_fmt = g.out_fmt = fmt;
  return gemm_bf16_launch(g, stream);
}
void gemm_timing_enable(int on);
int gemm_timing_read(double* total_ms, long long* launches);
int gemm_timing_get(long long index, double* ms, int* shape /* [10] */);

}  // namespace b2s
