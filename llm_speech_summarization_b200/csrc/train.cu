// train.cu -- training-mode forward (activations kept) and backward of the frozen LLM over packed sequences.
//
// The reference back-propagates total_loss through the frozen LlamaForCausalLM into the audio encoder
// (REF/trainer.py:62-64,317-322,373-374): only data gradients flow through the LLM (no weight gradients), and only
// through the student (audio-prompt) sequences -- the teacher pass runs under no_grad (REF/trainer.py:337).
// llama_forward_train is llama_prefill with every per-layer activation the backward needs written to caller-owned
// buffers (no recomputation, no copies: each layer writes its residual stream into a fresh buffer);
// llama_backward walks the layers in reverse with the same tcgen05 GEMM (transposed weight copies = dgrad),
// the attention backward kernels, and fused RMSNorm / SwiGLU / RoPE backward kernels.
#include "../../include/b2s.h"
#include "b2s_common.cuh"
#include "gemm_sm100.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

struct Carve {
  uint8_t* base;
  size_t off = 0, cap;
  Carve(void* b, size_t c) : base(reinterpret_cast<uint8_t*>(b)), cap(c) {}
  void* take(size_t bytes) {
    off = (off + 255) & ~static_cast<size_t>(255);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
};

GemmArgs lin(const void* A, const void* W, long long M, int N, int K) {
  GemmArgs g{};
  g.A = A;
  g.a_dim0 = K;
  g.a_row_stride = K;
  g.a_rows = static_cast<int>(M);
  g.W = W;
  g.w_rows = N;
  g.w_cols = K;
  g.M = static_cast<int>(M);
  g.N = N;
  g.batches = 1;
  g.groups = 1;
  g.taps = 1;
  g.k_per_tap = K;
  g.ldo = N;
  return g;
}

struct FwdWs {
  void *xn, *act, *xf;
  size_t bytes;
};
void plan_fwd(const b2s_llama_weights* w, long long rows, long long logit_rows, void* ws, size_t cap, FwdWs* p) {
  Carve c(ws, cap);
  p->xn = c.take(rows * w->hidden * 2);
  p->act = c.take(static_cast<size_t>(rows) * w->ffn * 2);
  p->xf = c.take(static_cast<size_t>(logit_rows > 0 ? logit_rows : 1) * w->hidden * 2);
  p->bytes = c.off + 256;
}

struct BwdWs {
  void *dh_bf16, *dact, *dgu, *dao, *dqkv;
  float *dxn, *delta, *dxf;
  size_t bytes;
};
void plan_bwd(const b2s_llama_weights* w, long long ms, long long n_dl, void* ws, size_t cap, BwdWs* p) {
  const size_t H = w->hidden, F = w->ffn, HqD = static_cast<size_t>(w->heads) * w->head_dim;
  const size_t qkv_cols = static_cast<size_t>(w->heads + 2 * w->kv_heads) * w->head_dim;
  Carve c(ws, cap);
  p->dh_bf16 = c.take(ms * H * 2);
  p->dact = c.take(ms * F * 2);
  p->dgu = c.take(ms * 2 * F * 2);
  p->dao = c.take(ms * HqD * 2);
  p->dqkv = c.take(ms * qkv_cols * 2);
  p->dxn = reinterpret_cast<float*>(c.take(ms * H * 4));
  p->delta = reinterpret_cast<float*>(c.take(ms * w->heads * 4));
  p->dxf = reinterpret_cast<float*>(c.take(static_cast<size_t>(n_dl > 0 ? n_dl : 1) * H * 4));
  p->bytes = c.off + 256;
}

}  // namespace

size_t llama_train_workspace_bytes(const b2s_llama_weights* w, int rows, int logit_rows) {
  if (!w || rows <= 0) return 0;
  FwdWs p;
  plan_fwd(w, rows, logit_rows, nullptr, 0, &p);
  return p.bytes;
}

size_t llama_backward_workspace_bytes(const b2s_llama_weights* w, int rows_bwd, int n_dl) {
  if (!w || rows_bwd <= 0) return 0;
  BwdWs p;
  plan_bwd(w, rows_bwd, n_dl, nullptr, 0, &p);
  return p.bytes;
}

int llama_forward_train(const b2s_llama_weights* w, const b2s_llama_saved* sv, int rows, const int* cu_seqlens,
                        int num_seqs, int max_seqlen, const int* positions, const int* logit_rows_index,
                        int logit_rows, void* logits_bf16, const int* tap_layers, int num_taps, const int* tap_rows_a,
                        const int* tap_rows_b, int pairs, float* fd_sq, void* workspace, size_t workspace_bytes,
                        cudaStream_t stream) {
  B2S_REQUIRE(w && sv && sv->h && sv->h_mid && sv->qkv && sv->ao && sv->lse && sv->gu && cu_seqlens && positions &&
                  workspace,
              "llama_forward_train: null pointer");
  B2S_REQUIRE(rows > 0 && num_seqs > 0 && max_seqlen > 0, "llama_forward_train: empty batch");
  B2S_REQUIRE(w->head_dim == 128 && w->hidden % 256 == 0 && w->ffn % 64 == 0, "llama_forward_train: unsupported shape");
  FwdWs pl;
  plan_fwd(w, rows, logit_rows, workspace, workspace_bytes, &pl);
  B2S_REQUIRE(pl.bytes <= workspace_bytes, "llama_forward_train: workspace too small (%zu < %zu)", workspace_bytes,
              pl.bytes);
  const int H = w->hidden, D = w->head_dim, Hq = w->heads, Hkv = w->kv_heads, F = w->ffn, Lyr = w->num_layers;
  const int fmt = w->fmt;
  const int qkv_cols = (Hq + 2 * Hkv) * D;
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  const size_t R = static_cast<size_t>(rows);
  int rc;
  for (int l = 0; l < Lyr; ++l) {
    float* h_in = sv->h + l * R * H;
    float* h_mid = sv->h_mid + l * R * H;
    float* h_out = sv->h + (l + 1) * R * H;
    __nv_bfloat16* qkv = reinterpret_cast<__nv_bfloat16*>(sv->qkv) + l * R * qkv_cols;
    __nv_bfloat16* ao = reinterpret_cast<__nv_bfloat16*>(sv->ao) + l * R * Hq * D;
    __nv_bfloat16* gu = reinterpret_cast<__nv_bfloat16*>(sv->gu) + l * R * 2 * F;
    float* lse = sv->lse + l * R * Hq;
    for (int t = 0; t < num_taps; ++t) {
      if (tap_layers[t] == l && pairs > 0) {
        rc = rowpair_sqdiff_fwd(h_in, tap_rows_a, tap_rows_b, fd_sq + static_cast<long long>(t) * pairs, pairs, H, stream);
        if (rc != B2S_OK) return rc;
      }
    }
    const b2s_llama_layer& L = w->layers[l];
    rc = rmsnorm_fwd(h_in, L.ln1_w, w->rms_eps, pl.xn, rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = lin(pl.xn, L.wqkv, rows, qkv_cols, H);
      g.epi = EPI_ROPE;
      g.out = qkv;
      g.rope_cs = w->rope_cs;
      g.positions = positions;
      g.rope_cols = (Hq + Hkv) * D;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    rc = attention_fwd(qkv, qkv + Hq * D, qkv + (Hq + Hkv) * D, qkv_cols, ao, Hq * D, cu_seqlens, num_seqs, max_seqlen,
                       rows, Hq, Hkv, D, scale, 1, lse, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = lin(ao, L.wo, rows, H, Hq * D);
      g.epi = EPI_RESID_F32;
      g.out = h_mid;
      g.resid = h_in;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    rc = rmsnorm_fwd(h_mid, L.ln2_w, w->rms_eps, pl.xn, rows, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {
      GemmArgs g = lin(pl.xn, L.wgu, rows, 2 * F, H);
      g.epi = EPI_SWIGLU;
      g.out = pl.act;
      g.ldo = F;
      g.out2 = gu;
      g.ld2 = 2 * F;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      GemmArgs g = lin(pl.act, L.wd, rows, H, F);
      g.epi = EPI_RESID_F32;
      g.out = h_out;
      g.resid = h_mid;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
  }
  if (logit_rows > 0) {
    B2S_REQUIRE(logit_rows_index && logits_bf16, "llama_forward_train: logits requested without buffers");
    rc = rmsnorm_gather_fwd(sv->h + Lyr * R * H, logit_rows_index, w->final_norm_w, w->rms_eps, pl.xf, logit_rows, H,
                            fmt,
                            stream);
    if (rc != B2S_OK) return rc;
    GemmArgs g = lin(pl.xf, w->lm_head, logit_rows, w->vocab, H);
    g.epi = EPI_BF16;
    g.out = logits_bf16;
    g.a_fmt = g.w_fmt = fmt;
    g.out_fmt = 0;  // logits stay bf16 (b2s.h)
    rc = gemm_bf16_launch(g, stream);
    if (rc != B2S_OK) return rc;
  }
  return B2S_OK;
}

int llama_backward(const b2s_llama_weights* w, const b2s_llama_weights_t* wt, const b2s_llama_saved* sv, int rows,
                   int rows_bwd, const int* cu_seqlens, int num_seqs_bwd, int max_seqlen, const void* d_logits,
                   const int* dl_rows_index, int n_dl, const int* tap_layers, int num_taps, const int* tap_rows_a,
                   const int* tap_rows_b, const float* tap_coef, const float* loss_scale, int pairs, float* dh,
                   void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  B2S_REQUIRE(w && wt && wt->layers && wt->lm_head_t && sv && cu_seqlens && d_logits && dl_rows_index && dh && workspace,
              "llama_backward: null pointer");
  B2S_REQUIRE(rows_bwd > 0 && rows_bwd <= rows && num_seqs_bwd > 0 && n_dl > 0, "llama_backward: bad sizes");
  BwdWs pl;
  plan_bwd(w, rows_bwd, n_dl, workspace, workspace_bytes, &pl);
  B2S_REQUIRE(pl.bytes <= workspace_bytes, "llama_backward: workspace too small (%zu < %zu)", workspace_bytes, pl.bytes);
  const int H = w->hidden, D = w->head_dim, Hq = w->heads, Hkv = w->kv_heads, F = w->ffn, Lyr = w->num_layers;
  const int fmt = w->fmt;
  const int qkv_cols = (Hq + 2 * Hkv) * D;
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  const size_t R = static_cast<size_t>(rows);
  const long long Ms = rows_bwd;
  int rc;

  B2S_CUDA_CHECK(cudaMemsetAsync(dh, 0, static_cast<size_t>(Ms) * H * 4, stream));
  B2S_CUDA_CHECK(cudaMemsetAsync(pl.dh_bf16, 0, static_cast<size_t>(Ms) * H * 2, stream));
  {  // LM head dgrad on the consumed rows: few output tiles and K = vocab, so split K over the whole chip
    B2S_CUDA_CHECK(cudaMemsetAsync(pl.dxf, 0, static_cast<size_t>(n_dl) * H * 4, stream));
    GemmArgs g = lin(d_logits, wt->lm_head_t, n_dl, H, w->vocab);
    g.epi = EPI_ACCUM_F32;
    g.out = pl.dxf;
    rc = gemm_launch_fmt(g, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  rc = rmsnorm_bwd(sv->h + Lyr * R * H, dl_rows_index, w->final_norm_w, w->rms_eps, pl.dxf, dh, dl_rows_index,
                   pl.dh_bf16, n_dl, H, fmt, stream);
  if (rc != B2S_OK) return rc;

  for (int l = Lyr - 1; l >= 0; --l) {
    const b2s_llama_layer& L = w->layers[l];
    const b2s_llama_layer_t& T = wt->layers[l];
    // feature-distillation gradient on hidden_states[l+1] (the input of layer l+1)
    for (int t = 0; t < num_taps; ++t) {
      if (tap_layers[t] == l + 1 && pairs > 0) {
        rc = add_rowdiff(sv->h + (l + 1) * R * H, tap_rows_a, tap_rows_b, tap_coef, loss_scale, dh, pl.dh_bf16, pairs, H,
                         fmt, stream);
        if (rc != B2S_OK) return rc;
      }
    }
    const __nv_bfloat16* qkv = reinterpret_cast<const __nv_bfloat16*>(sv->qkv) + l * R * qkv_cols;
    const __nv_bfloat16* ao = reinterpret_cast<const __nv_bfloat16*>(sv->ao) + l * R * Hq * D;
    const __nv_bfloat16* gu = reinterpret_cast<const __nv_bfloat16*>(sv->gu) + l * R * 2 * F;
    const float* lse = sv->lse + l * R * Hq;
    {  // down_proj dgrad
      GemmArgs g = lin(pl.dh_bf16, T.wd_t, Ms, F, H);
      g.epi = EPI_BF16;
      g.out = pl.dact;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    // (fusing this into the dgrad epilogue was measured: the GEMMs lose more than the 80 us launch saves)
    rc = swiglu_bwd(gu, pl.dact, pl.dgu, Ms, F, fmt, stream);
    if (rc != B2S_OK) return rc;
    {  // gate|up dgrad
      GemmArgs g = lin(pl.dgu, T.wgu_t, Ms, H, 2 * F);
      g.epi = EPI_F32;
      g.out = pl.dxn;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    rc = rmsnorm_bwd(sv->h_mid + l * R * H, nullptr, L.ln2_w, w->rms_eps, pl.dxn, dh, nullptr, pl.dh_bf16, Ms, H, fmt, stream);
    if (rc != B2S_OK) return rc;
    {  // o_proj dgrad
      GemmArgs g = lin(pl.dh_bf16, T.wo_t, Ms, Hq * D, H);
      g.epi = EPI_BF16;
      g.out = pl.dao;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {
      __nv_bfloat16* dqkv = reinterpret_cast<__nv_bfloat16*>(pl.dqkv);
      rc = attention_bwd(qkv, qkv + Hq * D, qkv + (Hq + Hkv) * D, qkv_cols, ao, Hq * D, pl.dao, Hq * D, lse, pl.delta,
                         dqkv, dqkv + Hq * D, dqkv + (Hq + Hkv) * D, qkv_cols, cu_seqlens, num_seqs_bwd, max_seqlen, Ms,
                         Hq, Hkv, D, scale, 1, w->rope_cs, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    {  // fused-QKV dgrad
      GemmArgs g = lin(pl.dqkv, T.wqkv_t, Ms, H, qkv_cols);
      g.epi = EPI_F32;
      g.out = pl.dxn;
      rc = gemm_launch_fmt(g, fmt, stream);
      if (rc != B2S_OK) return rc;
    }
    rc = rmsnorm_bwd(sv->h + l * R * H, nullptr, L.ln1_w, w->rms_eps, pl.dxn, dh, nullptr, pl.dh_bf16, Ms, H, fmt, stream);
    if (rc != B2S_OK) return rc;
  }
  return B2S_OK;
}

}  // namespace b2s
