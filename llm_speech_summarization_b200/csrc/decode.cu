// decode.cu -- greedy decode step with a KV cache: what HF `generate(inputs_embeds=...)`
// runs after the prefill in REF/inference.py:55-74 and REF/trainer.py:530-545 (LlamaModel.forward with
// past_key_values, TF/models/llama/modeling_llama.py:225-289,375-425).
//
// Cache layout: bf16 [layers][slots][2 * Hkv * D], row = k (post-RoPE) | v of one token; sequence b owns the slots
// [seq_start[b], seq_start[b] + capacity). The prefill scatters its K/V rows into the cache (kv_scatter), a decode step
// processes ONE new token per sequence:
//   h = embed[token] -> per layer { RMSNorm -> QKV GEMM (+RoPE at position = current length) -> append k|v ->
//   attention of the single query over the cached keys -> o GEMM -> RMSNorm -> gate|up GEMM (+SwiGLU) -> down GEMM }
//   -> final RMSNorm -> LM head.
// A decode step is weight-streaming bound (M = batch rows against 6.4 GB of weights): the GEMMs run with narrow
// tiles, and the two residual projections use the accumulate epilogue straight onto the fp32 residual stream so they
// can split K across the whole chip (out += partial IS the residual add).
#include "../../include/b2s.h"
#include "b2s_common.cuh"
#include "gemm_sm100.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

// cache[slot(row)] = qkv[row, col0 : col0 + width]
__global__ void __launch_bounds__(256)
kv_scatter_kernel(const __nv_bfloat16* __restrict__ qkv, long long ld, int col0, int width,
                  __nv_bfloat16* __restrict__ cache, const int* __restrict__ slot_of_row, const int* __restrict__ base,
                  const int* __restrict__ len, long long rows) {
  const int per_row = width / 8;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= rows * per_row) return;
  const long long r = i / per_row;
  const int c = static_cast<int>(i - r * per_row) * 8;
  const int slot = slot_of_row ? slot_of_row[r] : base[r] + len[r];
  if (slot < 0) return;
  *reinterpret_cast<uint4*>(cache + static_cast<long long>(slot) * width + c) =
      *reinterpret_cast<const uint4*>(qkv + r * ld + col0 + c);
}

// one query row per (sequence, head) against the cached keys: 4 warps split the keys, lane = 4 of the 128 dims
constexpr int kDecWarps = 4;
__global__ void __launch_bounds__(kDecWarps * 32)
decode_attn_kernel(const __nv_bfloat16* __restrict__ q, long long ldq, const __nv_bfloat16* __restrict__ cache,
                   const int* __restrict__ seq_start, const int* __restrict__ seq_len, __nv_bfloat16* __restrict__ o,
                   int Hq, int Hkv, float scale_log2) {
  constexpr int D = 128;
  __shared__ float s_m[kDecWarps], s_l[kDecWarps], s_acc[kDecWarps][D];
  const int b = blockIdx.x, h = blockIdx.y;
  const int kvh = h / (Hq / Hkv);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = seq_len[b] + 1;  // the new token's k|v were appended before this launch
  const long long ldkv = 2LL * Hkv * D;
  const __nv_bfloat16* kbase = cache + static_cast<long long>(seq_start[b]) * ldkv + kvh * D + lane * 4;
  const __nv_bfloat16* vbase = kbase + Hkv * D;
  float qv[4];
  {
    const uint2 u = *reinterpret_cast<const uint2*>(q + b * ldq + h * D + lane * 4);
    qv[0] = bf16_lo(u.x) * scale_log2; qv[1] = bf16_hi(u.x) * scale_log2;
    qv[2] = bf16_lo(u.y) * scale_log2; qv[3] = bf16_hi(u.y) * scale_log2;
  }
  float m = -INFINITY, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = warp; j < n; j += kDecWarps) {
    const uint2 ku = *reinterpret_cast<const uint2*>(kbase + j * ldkv);
    float s = qv[0] * bf16_lo(ku.x) + qv[1] * bf16_hi(ku.x) + qv[2] * bf16_lo(ku.y) + qv[3] * bf16_hi(ku.y);
    s = warp_sum(s);
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    const uint2 vu = *reinterpret_cast<const uint2*>(vbase + j * ldkv);
    l = l * corr + p;
    acc[0] = acc[0] * corr + p * bf16_lo(vu.x);
    acc[1] = acc[1] * corr + p * bf16_hi(vu.x);
    acc[2] = acc[2] * corr + p * bf16_lo(vu.y);
    acc[3] = acc[3] * corr + p * bf16_hi(vu.y);
    m = mn;
  }
  if (lane == 0) {
    s_m[warp] = m;
    s_l[warp] = l;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) s_acc[warp][lane * 4 + i] = acc[i];
  __syncthreads();
  if (warp == 0) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) mm = fmaxf(mm, s_m[w]);
    float ll = 0.f, out[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) {
      const float c = s_m[w] == -INFINITY ? 0.f : exp2f(s_m[w] - mm);
      ll += s_l[w] * c;
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] += s_acc[w][lane * 4 + i] * c;
    }
    const float inv = 1.0f / ll;
    uint2 u;
    u.x = pack_bf16(out[0] * inv, out[1] * inv);
    u.y = pack_bf16(out[2] * inv, out[3] * inv);
    *reinterpret_cast<uint2*>(o + static_cast<long long>(b) * Hq * D + h * D + lane * 4) = u;
  }
}

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

struct DecWs {
  float* h;
  void *xn, *qkv, *ao, *act;
  size_t bytes;
};
void plan_dec(const b2s_llama_weights* w, int batch, void* ws, size_t cap, DecWs* p) {
  const size_t B = batch, H = w->hidden, D = w->head_dim;
  Carve c(ws, cap);
  p->h = reinterpret_cast<float*>(c.take(B * H * 4));
  p->xn = c.take(B * H * 2 + 4096);
  p->qkv = c.take(B * (w->heads + 2 * w->kv_heads) * D * 2 + 4096);
  p->ao = c.take(B * w->heads * D * 2 + 4096);
  p->act = c.take(B * w->ffn * 2 + 4096);
  p->bytes = c.off + 256;
}

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

#define RC(expr)                   \
  do {                             \
    int _rc = (expr);              \
    if (_rc != B2S_OK) return _rc; \
  } while (0)

}  // namespace

size_t llama_kv_cache_bytes(const b2s_llama_weights* w, int slots) {
  if (w == nullptr || slots <= 0) return 0;
  return static_cast<size_t>(w->num_layers) * slots * 2 * w->kv_heads * w->head_dim * 2;
}

// prefill side: copy the k|v columns of the packed QKV buffer of layer `layer` into the cache
int kv_cache_store(const b2s_llama_weights* w, int layer, const void* qkv, long long rows, void* kv_cache, int kv_slots,
                   const int* slot_of_row, cudaStream_t stream) {
  const int D = w->head_dim, width = 2 * w->kv_heads * D;
  const long long n = rows * (width / 8);
  __nv_bfloat16* cache = reinterpret_cast<__nv_bfloat16*>(kv_cache) + static_cast<size_t>(layer) * kv_slots * width;
  kv_scatter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(qkv), static_cast<long long>(w->heads + 2 * w->kv_heads) * D, w->heads * D,
      width, cache, slot_of_row, nullptr, nullptr, rows);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

size_t llama_decode_workspace_bytes(const b2s_llama_weights* w, int batch) {
  if (w == nullptr || batch <= 0) return 0;
  DecWs p;
  plan_dec(w, batch, nullptr, 0, &p);
  return p.bytes;
}

int llama_decode_step(const b2s_llama_weights* w, const void* embed_table, const int* token_ids, int batch,
                      void* kv_cache, int kv_slots, const int* seq_start, const int* seq_len, void* logits_bf16,
                      void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  B2S_REQUIRE(w && embed_table && token_ids && kv_cache && seq_start && seq_len && logits_bf16 && workspace,
              "llama_decode_step: null pointer");
  B2S_REQUIRE(batch > 0 && kv_slots > 0, "llama_decode_step: empty batch");
  B2S_REQUIRE(w->head_dim == 128, "llama_decode_step: head_dim must be 128 (got %d)", w->head_dim);
  DecWs p;
  plan_dec(w, batch, workspace, workspace_bytes, &p);
  B2S_REQUIRE(p.bytes <= workspace_bytes, "llama_decode_step: workspace too small: need %zu bytes, got %zu", p.bytes,
              workspace_bytes);
  const int B = batch, H = w->hidden, D = w->head_dim, Hq = w->heads, Hkv = w->kv_heads, F = w->ffn;
  const int qkv_cols = (Hq + 2 * Hkv) * D, width = 2 * Hkv * D;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(D));

  RC(embed_splice_fwd(embed_table, nullptr, token_ids, p.h, B, H, stream));
  for (int l = 0; l < w->num_layers; ++l) {
    const b2s_llama_layer& L = w->layers[l];
    __nv_bfloat16* cache = reinterpret_cast<__nv_bfloat16*>(kv_cache) + static_cast<size_t>(l) * kv_slots * width;
    RC(rmsnorm_fwd(p.h, L.ln1_w, w->rms_eps, p.xn, B, H, stream));
    {
      GemmArgs g = lin(p.xn, L.wqkv, B, qkv_cols, H);
      g.epi = EPI_ROPE;
      g.out = p.qkv;
      g.rope_cs = w->rope_cs;
      g.positions = seq_len;  // position of the new token = current length of its sequence
      g.rope_cols = (Hq + Hkv) * D;
      g.block_n = 128;
      g.cta_group = 1;
      RC(gemm_bf16_launch(g, stream));
    }
    {
      const long long n = static_cast<long long>(B) * (width / 8);
      kv_scatter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(p.qkv), qkv_cols, Hq * D, width, cache, nullptr, seq_start, seq_len, B);
      B2S_LAUNCH_CHECK();
      decode_attn_kernel<<<dim3(B, Hq), kDecWarps * 32, 0, stream>>>(
          reinterpret_cast<const __nv_bfloat16*>(p.qkv), qkv_cols, cache, seq_start, seq_len,
          reinterpret_cast<__nv_bfloat16*>(p.ao), Hq, Hkv, scale_log2);
      B2S_LAUNCH_CHECK();
    }
    {
      GemmArgs g = lin(p.ao, L.wo, B, H, Hq * D);
      g.epi = EPI_ACCUM_F32;  // h += ao . Wo^T, split-K over the whole chip
      g.out = p.h;
      g.cta_group = 1;
      RC(gemm_bf16_launch(g, stream));
    }
    RC(rmsnorm_fwd(p.h, L.ln2_w, w->rms_eps, p.xn, B, H, stream));
    {
      GemmArgs g = lin(p.xn, L.wgu, B, 2 * F, H);
      g.epi = EPI_SWIGLU;
      g.out = p.act;
      g.ldo = F;
      g.block_n = 128;
      g.cta_group = 1;
      RC(gemm_bf16_launch(g, stream));
    }
    {
      GemmArgs g = lin(p.act, L.wd, B, H, F);
      g.epi = EPI_ACCUM_F32;
      g.out = p.h;
      g.cta_group = 1;
      RC(gemm_bf16_launch(g, stream));
    }
  }
  RC(rmsnorm_fwd(p.h, w->final_norm_w, w->rms_eps, p.xn, B, H, stream));
  {
    GemmArgs g = lin(p.xn, w->lm_head, B, w->vocab, H);
    g.epi = EPI_BF16;
    g.out = logits_bf16;
    g.cta_group = 1;
    RC(gemm_bf16_launch(g, stream));
  }
  return B2S_OK;
}

}  // namespace b2s
