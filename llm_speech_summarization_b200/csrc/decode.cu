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
// Measured and rejected in round 2 (profiles/r02_decode.md): the whole step as ONE persistent cooperative kernel (one
// CTA per SM, equal contiguous row ranges per CTA, grid barriers between the five phases of a layer, next-phase weight
// prefetch): parity-green, but 2.75 ms / token against 2.39 ms for the PDL-chained launches below at batch 1 -- with
// fp16 operands the per-element convert + FMA work, not the launch gaps, is what keeps the step at ~0.42 of HBM peak.
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

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// one query row per (sequence, head) against the cached keys plus the new token's own k|v (read from the QKV buffer and
// appended to the cache by the first query head of each kv group): 4 warps split the keys, lane = 4 of the 128 dims
constexpr int kDecWarps = 4;
__global__ void __launch_bounds__(kDecWarps * 32)
decode_attn_kernel(const __nv_bfloat16* __restrict__ qkv, long long ldq, __nv_bfloat16* __restrict__ cache,
                   const int* __restrict__ seq_start, const int* __restrict__ seq_len, __nv_bfloat16* __restrict__ o,
                   int Hq, int Hkv, float scale_log2, int f16) {
  constexpr int D = 128;
  __shared__ float s_m[kDecWarps], s_l[kDecWarps], s_acc[kDecWarps][D];
  pdl_launch_dependents();
  const int b = blockIdx.x, h = blockIdx.y;
  const int group = Hq / Hkv, kvh = h / group;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait();  // everything below depends on the QKV projection of this step
  const int n = seq_len[b];  // cached tokens; the new token is key number n
  const long long ldkv = 2LL * Hkv * D;
  __nv_bfloat16* kbase = cache + static_cast<long long>(seq_start[b]) * ldkv + kvh * D + lane * 4;
  const __nv_bfloat16* vbase = kbase + Hkv * D;
  const __nv_bfloat16* qrow = qkv + b * ldq;
  const uint2 knew = *reinterpret_cast<const uint2*>(qrow + (Hq + kvh) * D + lane * 4);
  const uint2 vnew = *reinterpret_cast<const uint2*>(qrow + (Hq + Hkv + kvh) * D + lane * 4);
  if (warp == 0 && h % group == 0) {  // append: nobody reads slot n of the cache during this launch
    *reinterpret_cast<uint2*>(kbase + n * ldkv) = knew;
    *reinterpret_cast<uint2*>(kbase + Hkv * D + n * ldkv) = vnew;
  }
  float qv[4];
  {
    const uint2 u = *reinterpret_cast<const uint2*>(qrow + h * D + lane * 4);
    const float2 a = unpack_h16(u.x, f16), b2 = unpack_h16(u.y, f16);
    qv[0] = a.x * scale_log2; qv[1] = a.y * scale_log2;
    qv[2] = b2.x * scale_log2; qv[3] = b2.y * scale_log2;
  }
  float m = -INFINITY, l = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int j = warp; j <= n; j += kDecWarps) {
    const uint2 ku = j < n ? *reinterpret_cast<const uint2*>(kbase + j * ldkv) : knew;
    const uint2 vu = j < n ? *reinterpret_cast<const uint2*>(vbase + j * ldkv) : vnew;
    const float2 k0 = unpack_h16(ku.x, f16), k1 = unpack_h16(ku.y, f16);
    const float2 v0 = unpack_h16(vu.x, f16), v1 = unpack_h16(vu.y, f16);
    float s = qv[0] * k0.x + qv[1] * k0.y + qv[2] * k1.x + qv[3] * k1.y;
    s = warp_sum(s);
    const float mn = fmaxf(m, s);
    const float corr = exp2f(m - mn), p = exp2f(s - mn);
    l = l * corr + p;
    acc[0] = acc[0] * corr + p * v0.x;
    acc[1] = acc[1] * corr + p * v0.y;
    acc[2] = acc[2] * corr + p * v1.x;
    acc[3] = acc[3] * corr + p * v1.y;
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
    u.x = pack_h16(out[0] * inv, out[1] * inv, f16);
    u.y = pack_h16(out[2] * inv, out[3] * inv, f16);
    *reinterpret_cast<uint2*>(o + static_cast<long long>(b) * Hq * D + h * D + lane * 4) = u;
  }
}

// ---- weight-streaming GEMV for decode batches of <= 4 rows -------------------------------------------------------
// out[b, n] = epi( sum_k x[b, k] * W[n, k] ): every weight element is read exactly once with 16-byte loads (the step is
// HBM-bound: 2 bytes of weight per B FMAs), the <= 4 activation rows live in shared memory, one warp owns one output
// column -- or, for the paired epilogues, the two columns a SwiGLU / rotate-half pair needs (n and n + 64 of a
// 128-row block), so the epilogue needs no second pass. No split-K: each output element has one owner, the fp32
// residual add is a plain read-modify-write.
enum GemvEpi : int { GV_BF16 = 0, GV_ACCUM_F32 = 1, GV_SWIGLU = 2, GV_ROPE = 3 };
constexpr int kGvWarps = 8;

__device__ __forceinline__ float dot8_f16(const uint4& w, const uint4& x) {
  const float2 w0 = unpack_f16(w.x), w1 = unpack_f16(w.y), w2 = unpack_f16(w.z), w3 = unpack_f16(w.w);
  const float2 x0 = unpack_f16(x.x), x1 = unpack_f16(x.y), x2 = unpack_f16(x.z), x3 = unpack_f16(x.w);
  float s = w0.x * x0.x;
  s = fmaf(w0.y, x0.y, s);
  s = fmaf(w1.x, x1.x, s);
  s = fmaf(w1.y, x1.y, s);
  s = fmaf(w2.x, x2.x, s);
  s = fmaf(w2.y, x2.y, s);
  s = fmaf(w3.x, x3.x, s);
  s = fmaf(w3.y, x3.y, s);
  return s;
}
__device__ __forceinline__ float dot8_bf16(const uint4& w, const uint4& x) {
  float s = bf16_lo(w.x) * bf16_lo(x.x);
  s = fmaf(bf16_hi(w.x), bf16_hi(x.x), s);
  s = fmaf(bf16_lo(w.y), bf16_lo(x.y), s);
  s = fmaf(bf16_hi(w.y), bf16_hi(x.y), s);
  s = fmaf(bf16_lo(w.z), bf16_lo(x.z), s);
  s = fmaf(bf16_hi(w.z), bf16_hi(x.z), s);
  s = fmaf(bf16_lo(w.w), bf16_lo(x.w), s);
  s = fmaf(bf16_hi(w.w), bf16_hi(x.w), s);
  return s;
}

// F16: operand format of x / W (compile-time: the dot product is the hot loop); out_f16: format of 16-bit outputs
template <int NB, bool PAIRED, bool F16>
__global__ void __launch_bounds__(kGvWarps * 32, 2)
gemv_h16_kernel(const void* __restrict__ x, const float* __restrict__ norm_w, float norm_eps,
                const __nv_bfloat16* __restrict__ W, int B, int N, int K, int epi, void* out, long long ldo,
                const float* __restrict__ rope_cs, const int* __restrict__ positions, int rope_cols, int out_f16) {
  auto dot8 = [](const uint4& w, const uint4& xv) { return F16 ? dot8_f16(w, xv) : dot8_bf16(w, xv); };
  extern __shared__ uint4 sx[];  // [NB][K / 8] bf16 activations
  __shared__ float s_part[NB][kGvWarps];
  pdl_launch_dependents();  // the next kernel may start prefetching ITS weights while this one runs
  const int k8 = K / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int units = PAIRED ? N / 2 : N;
  int u = blockIdx.x * kGvWarps + warp;
  // weights do not depend on the previous kernel: get the first chunks in flight before waiting for it
  constexpr int kPre = 4;
  uint4 pre0[kPre], pre1[kPre];
  {
    const int n0 = u < units ? (PAIRED ? (u >> 6) * 128 + (u & 63) : u) : 0;
    const uint4* w0 = reinterpret_cast<const uint4*>(W + static_cast<long long>(n0) * K);
#pragma unroll
    for (int i = 0; i < kPre; ++i) {
      const int c = lane + 32 * i;
      pre0[i] = (u < units && c < k8) ? ld_stream_u4(w0 + c) : make_uint4(0u, 0u, 0u, 0u);
      pre1[i] = (PAIRED && u < units && c < k8) ? ld_stream_u4(w0 + 64 * k8 + c) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  pdl_wait();
  if (norm_w == nullptr) {  // x: bf16 [B, K]
    for (int i = threadIdx.x; i < NB * k8; i += blockDim.x) {
      const int b = i / k8;
      sx[i] = b < B ? reinterpret_cast<const uint4*>(x)[static_cast<long long>(b) * k8 + (i - b * k8)]
                    : make_uint4(0u, 0u, 0u, 0u);
    }
  } else {  // x: fp32 residual stream [B, K]; stage RMSNorm(x) * w rounded to bf16 (LlamaRMSNorm)
    const float* xf = reinterpret_cast<const float*>(x);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float q = 0.f;
      if (b < B)
        for (int c = threadIdx.x; c < k8; c += blockDim.x) {
          float v[8];
          ld8f(xf + static_cast<long long>(b) * K + c * 8, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) q = fmaf(v[j], v[j], q);
        }
      q = warp_sum(q);
      if (lane == 0) s_part[b][warp] = q;
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float q = 0.f;
#pragma unroll
      for (int w = 0; w < kGvWarps; ++w) q += s_part[b][w];
      const float rstd = rsqrtf(q / static_cast<float>(K) + norm_eps);
      for (int c = threadIdx.x; c < k8; c += blockDim.x) {
        uint4 pk = make_uint4(0u, 0u, 0u, 0u);
        if (b < B) {
          float v[8], wt[8];
          ld8f(xf + static_cast<long long>(b) * K + c * 8, v);
          ld8f(norm_w + c * 8, wt);
          pk.x = pack_h16(wt[0] * (v[0] * rstd), wt[1] * (v[1] * rstd), F16);
          pk.y = pack_h16(wt[2] * (v[2] * rstd), wt[3] * (v[3] * rstd), F16);
          pk.z = pack_h16(wt[4] * (v[4] * rstd), wt[5] * (v[5] * rstd), F16);
          pk.w = pack_h16(wt[6] * (v[6] * rstd), wt[7] * (v[7] * rstd), F16);
        }
        sx[b * k8 + c] = pk;
      }
    }
  }
  __syncthreads();
  bool first = true;
  for (; u < units; u += gridDim.x * kGvWarps, first = false) {
    const int n0 = PAIRED ? (u >> 6) * 128 + (u & 63) : u;
    const uint4* w0 = reinterpret_cast<const uint4*>(W + static_cast<long long>(n0) * K);
    const uint4* w1 = w0 + 64 * k8;  // row n0 + 64
    float a0[NB], a1[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) a0[b] = a1[b] = 0.f;
    int c = lane;
    if (first) {  // consume the prefetched chunks
#pragma unroll
      for (int i = 0; i < kPre; ++i, c += 32) {
        if (c < k8) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const uint4 xv = sx[b * k8 + c];
            a0[b] += dot8(pre0[i], xv);
            if (PAIRED) a1[b] += dot8(pre1[i], xv);
          }
        }
      }
    }
#pragma unroll 4
    for (; c < k8; c += 32) {
      const uint4 wv0 = ld_stream_u4(w0 + c);
      uint4 wv1 = make_uint4(0u, 0u, 0u, 0u);
      if (PAIRED) wv1 = ld_stream_u4(w1 + c);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint4 xv = sx[b * k8 + c];
        a0[b] += dot8(wv0, xv);
        if (PAIRED) a1[b] += dot8(wv1, xv);
      }
    }
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const float s0 = warp_sum(a0[b]);
      const float s1 = PAIRED ? warp_sum(a1[b]) : 0.f;
      if (lane == b) {
        v0 = s0;
        v1 = s1;
      }
    }
    if (lane < B) {
      const long long row = static_cast<long long>(lane) * ldo;
      if (epi == GV_BF16) {
        reinterpret_cast<uint16_t*>(out)[row + n0] = float_to_h16(v0, out_f16);
      } else if (epi == GV_ACCUM_F32) {
        reinterpret_cast<float*>(out)[row + n0] += v0;
      } else if (epi == GV_SWIGLU) {
        reinterpret_cast<uint16_t*>(out)[row + (u >> 6) * 64 + (u & 63)] = float_to_h16(silu(v0) * v1, out_f16);
      } else {  // GV_ROPE: rotate-half pair (d, d + 64) of a 128-wide head
        float lo = v0, hi = v1;
        if (n0 < rope_cols) {
          const float* cs = rope_cs + static_cast<long long>(positions[lane]) * 128 + (u & 63);
          const float cc = cs[0], sn = cs[64];
          lo = v0 * cc - v1 * sn;
          hi = v1 * cc + v0 * sn;
        }
        reinterpret_cast<uint16_t*>(out)[row + n0] = float_to_h16(lo, out_f16);
        reinterpret_cast<uint16_t*>(out)[row + n0 + 64] = float_to_h16(hi, out_f16);
      }
    }
  }
}

// launch with programmatic dependent launch: the kernel's prologue (weight prefetch) overlaps its predecessor's tail
template <typename Kern, typename... Args>
int launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  B2S_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launch();
  return B2S_OK;
}

template <int NB, bool PAIRED, bool F16>
int launch_gemv(const void* x, const float* norm_w, float norm_eps, const void* W, int B, int N, int K, int epi,
                void* out, long long ldo, const float* rope_cs, const int* positions, int rope_cols, int out_f16,
                cudaStream_t stream) {
  auto kern = gemv_h16_kernel<NB, PAIRED, F16>;
  const int smem = NB * K * 2;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  const int units = PAIRED ? N / 2 : N;
  int blocks = (units + kGvWarps - 1) / kGvWarps;
  const int cap = 4 * num_sms();
  if (blocks > cap) blocks = cap;
  return launch_pdl(kern, dim3(blocks), dim3(kGvWarps * 32), static_cast<size_t>(smem), stream, x, norm_w, norm_eps,
                    reinterpret_cast<const __nv_bfloat16*>(W), B, N, K, epi, out, ldo, rope_cs, positions, rope_cols, out_f16);
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

// x: 16-bit [B <= 4, K] in format fmt (norm_w == nullptr) or the fp32 residual stream [B, K] with RMSNorm(norm_w, eps)
// fused into the staging; W 16-bit [N, K] in fmt; 16-bit outputs in out_fmt; see GemvEpi
int gemv_h16(const void* x, const float* norm_w, float norm_eps, const void* W, int B, int N, int K, int epi, void* out,
             long long ldo, const float* rope_cs, const int* positions, int rope_cols, int fmt, int out_fmt,
             cudaStream_t stream) {
  B2S_REQUIRE(x && W && out && B >= 1 && B <= 4, "gemv: needs 1..4 rows");
  B2S_REQUIRE(K % 8 == 0 && K * 2 * 4 <= 200 * 1024, "gemv: K must be a multiple of 8 and fit shared memory");
  const bool paired = epi == GV_SWIGLU || epi == GV_ROPE;
  if (paired) B2S_REQUIRE(N % 128 == 0, "gemv: paired epilogues need N %% 128 == 0");
  if (epi == GV_ROPE) B2S_REQUIRE(rope_cs && positions, "gemv: rope epilogue needs tables");
#define B2S_GEMV_ARGS x, norm_w, norm_eps, W, B, N, K, epi, out, ldo, rope_cs, positions, rope_cols, out_fmt, stream
#define B2S_GEMV(NB)                                                                                    \
  if (fmt != 0) return paired ? launch_gemv<NB, true, true>(B2S_GEMV_ARGS) : launch_gemv<NB, false, true>(B2S_GEMV_ARGS); \
  return paired ? launch_gemv<NB, true, false>(B2S_GEMV_ARGS) : launch_gemv<NB, false, false>(B2S_GEMV_ARGS)
  if (B == 1) { B2S_GEMV(1); }
  if (B == 2) { B2S_GEMV(2); }
  B2S_GEMV(4);
#undef B2S_GEMV
#undef B2S_GEMV_ARGS
}

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
  const int B = batch, H = w->hidden, D = w->head_dim, Hq = w->heads, Hkv = w->kv_heads, F = w->ffn, fmt = w->fmt;
  const int qkv_cols = (Hq + 2 * Hkv) * D, width = 2 * Hkv * D;
  const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(D));

  RC(embed_splice_fwd(embed_table, nullptr, token_ids, p.h, B, H, fmt, stream));
  const bool gv = B <= 4;  // weight-streaming GEMV path (RMSNorm fused into its staging, PDL-chained launches)
  for (int l = 0; l < w->num_layers; ++l) {
    const b2s_llama_layer& L = w->layers[l];
    __nv_bfloat16* cache = reinterpret_cast<__nv_bfloat16*>(kv_cache) + static_cast<size_t>(l) * kv_slots * width;
    if (gv) {
      RC(gemv_h16(p.h, L.ln1_w, w->rms_eps, L.wqkv, B, qkv_cols, H, GV_ROPE, p.qkv, qkv_cols, w->rope_cs, seq_len,
                  (Hq + Hkv) * D, fmt, fmt, stream));
      RC(launch_pdl(decode_attn_kernel, dim3(B, Hq), dim3(kDecWarps * 32), 0, stream,
                    reinterpret_cast<const __nv_bfloat16*>(p.qkv), static_cast<long long>(qkv_cols), cache, seq_start,
                    seq_len, reinterpret_cast<__nv_bfloat16*>(p.ao), Hq, Hkv, scale_log2, fmt));
      RC(gemv_h16(p.ao, nullptr, 0.f, L.wo, B, H, Hq * D, GV_ACCUM_F32, p.h, H, nullptr, nullptr, 0, fmt, fmt, stream));
      RC(gemv_h16(p.h, L.ln2_w, w->rms_eps, L.wgu, B, 2 * F, H, GV_SWIGLU, p.act, F, nullptr, nullptr, 0, fmt, fmt,
                  stream));
      RC(gemv_h16(p.act, nullptr, 0.f, L.wd, B, H, F, GV_ACCUM_F32, p.h, H, nullptr, nullptr, 0, fmt, fmt, stream));
      continue;
    }
    RC(rmsnorm_fwd(p.h, L.ln1_w, w->rms_eps, p.xn, B, H, fmt, stream));
    {
      GemmArgs g = lin(p.xn, L.wqkv, B, qkv_cols, H);
      g.epi = EPI_ROPE;
      g.out = p.qkv;
      g.rope_cs = w->rope_cs;
      g.positions = seq_len;  // position of the new token = current length of its sequence
      g.rope_cols = (Hq + Hkv) * D;
      g.block_n = 128;
      g.cta_group = 1;
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    decode_attn_kernel<<<dim3(B, Hq), kDecWarps * 32, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(p.qkv), qkv_cols, cache, seq_start, seq_len,
        reinterpret_cast<__nv_bfloat16*>(p.ao), Hq, Hkv, scale_log2, fmt);
    B2S_LAUNCH_CHECK();
    {
      GemmArgs g = lin(p.ao, L.wo, B, H, Hq * D);
      g.epi = EPI_ACCUM_F32;  // h += ao . Wo^T, split-K over the whole chip
      g.out = p.h;
      g.cta_group = 1;
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    RC(rmsnorm_fwd(p.h, L.ln2_w, w->rms_eps, p.xn, B, H, fmt, stream));
    {
      GemmArgs g = lin(p.xn, L.wgu, B, 2 * F, H);
      g.epi = EPI_SWIGLU;
      g.out = p.act;
      g.ldo = F;
      g.block_n = 128;
      g.cta_group = 1;
      RC(gemm_launch_fmt(g, fmt, stream));
    }
    {
      GemmArgs g = lin(p.act, L.wd, B, H, F);
      g.epi = EPI_ACCUM_F32;
      g.out = p.h;
      g.cta_group = 1;
      RC(gemm_launch_fmt(g, fmt, stream));
    }
  }
  if (gv) {
    RC(gemv_h16(p.h, w->final_norm_w, w->rms_eps, w->lm_head, B, w->vocab, H, GV_BF16, logits_bf16, w->vocab, nullptr,
                nullptr, 0, fmt, 0 /* logits stay bf16 */, stream));
  } else {
    RC(rmsnorm_fwd(p.h, w->final_norm_w, w->rms_eps, p.xn, B, H, fmt, stream));
    GemmArgs g = lin(p.xn, w->lm_head, B, w->vocab, H);
    g.epi = EPI_BF16;
    g.out = logits_bf16;
    g.cta_group = 1;
    g.a_fmt = g.w_fmt = fmt;
    g.out_fmt = 0;  // logits stay bf16
    RC(gemm_bf16_launch(g, stream));
  }
  return B2S_OK;
}

}  // namespace b2s
