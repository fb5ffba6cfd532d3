// attention.cu -- packed variable-length flash attention (forward), bf16 in / fp32 online softmax / bf16 out.
//
// Covers both attention shapes on the hot path:
//   * HuBERT / Whisper encoder self-attention: non-causal, 16 heads x 64, T <= 1500, no mask
//     (TF/models/hubert/modeling_hubert.py:262-345; attention_mask is None on this path);
//   * Llama / MiniChat prefill: causal, GQA 24q/8kv (or MHA 24/24) x 128, one independent causal mask per
//     packed utterance (TF/models/llama/modeling_llama.py:225-289) -- numerically equivalent to the reference's
//     left-pad + attention-mask batching (REF/utils.py:136-146) because padded keys carry zero probability.
//
// Round-1 implementation: FA2-style tiling (64 queries x 64 keys per step, 4 warps x 16 query rows) on
// mma.sync.m16n8k16 with cp.async double-buffered K/V and an XOR-swizzled shared layout. Attention is ~1.5 %
// of the path's FLOPs; the tcgen05/TMEM version is scheduled after the GEMMs.
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kBM = 64;
constexpr int kBN = 64;
constexpr int kAttnThreads = 128;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 => zero-fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int D>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * (D * 2) + ((chunk ^ (row & 7)) << 4));
}

// copy `rows_valid` rows of D bf16 (global row stride ld) into a swizzled [64][D] tile, zero-filling the rest
template <int D>
__device__ __forceinline__ void load_tile(uint32_t smem, const __nv_bfloat16* g, long long ld, int rows_valid) {
  constexpr int CH = D / 8;
  for (int i = threadIdx.x; i < 64 * CH; i += kAttnThreads) {
    const int r = i / CH, c = i - r * CH;
    const bool ok = r < rows_valid;
    cp_async16(smem + swz<D>(r, c), ok ? static_cast<const void*>(g + r * ld + c * 8) : static_cast<const void*>(g),
               ok);
  }
}

template <int D>
__global__ void __launch_bounds__(kAttnThreads)
attn_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                const __nv_bfloat16* __restrict__ v, long long ld, __nv_bfloat16* __restrict__ o, long long ldo,
                const int* __restrict__ cu, int Hq, int Hkv, float scale_log2, int causal) {
  constexpr int ROWB = D * 2;
  constexpr int KSTEPS = D / 16;
  constexpr int DBLK = D / 8;
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t sK = sQ + kBM * ROWB;
  const uint32_t sV = sK + 2 * kBN * ROWB;

  const int seq = blockIdx.z, h = blockIdx.y;
  const int s0 = cu[seq];
  const int L = cu[seq + 1] - s0;
  const int q0 = blockIdx.x * kBM;
  if (q0 >= L) return;
  const int hk = h / (Hq / Hkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* qg = q + static_cast<long long>(s0 + q0) * ld + h * D;
  const __nv_bfloat16* kg = k + static_cast<long long>(s0) * ld + hk * D;
  const __nv_bfloat16* vg = v + static_cast<long long>(s0) * ld + hk * D;

  const int kv_len = causal ? min(L, q0 + kBM) : L;
  const int nblk = (kv_len + kBN - 1) / kBN;

  load_tile<D>(sQ, qg, ld, min(kBM, L - q0));
  load_tile<D>(sK, kg, ld, min(kBN, L));
  load_tile<D>(sV, vg, ld, min(kBN, L));
  cp_async_commit();

  uint32_t qf[KSTEPS][4];
  float oacc[DBLK][4];
#pragma unroll
  for (int i = 0; i < DBLK; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int qrow0 = q0 + warp * 16 + g;  // this thread's two query rows: qrow0, qrow0 + 8 (sequence-local)

  for (int j = 0; j < nblk; ++j) {
    const int st = j & 1;
    if (j + 1 < nblk) {
      const int kn = (j + 1) * kBN;
      load_tile<D>(sK + (st ^ 1) * kBN * ROWB, kg + static_cast<long long>(kn) * ld, ld, min(kBN, L - kn));
      load_tile<D>(sV + (st ^ 1) * kBN * ROWB, vg + static_cast<long long>(kn) * ld, ld, min(kBN, L - kn));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = ks * 2 + (lane >> 4);
        ldmatrix_x4(sQ + swz<D>(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }

    // S = Q K^T  (16 x 64 per warp)
    float sacc[kBN / 8][4];
#pragma unroll
    for (int i = 0; i < kBN / 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) sacc[i][jj] = 0.f;
    const uint32_t sKs = sK + st * kBN * ROWB;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
      for (int nb = 0; nb < kBN / 16; ++nb) {
        uint32_t b0, b1, b2, b3;
        const int row = nb * 16 + (lane & 7) + (lane >> 4) * 8;
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        ldmatrix_x4(sKs + swz<D>(row, chunk), b0, b1, b2, b3);
        mma_bf16(sacc[2 * nb], qf[ks], b0, b1);
        mma_bf16(sacc[2 * nb + 1], qf[ks], b2, b3);
      }
    }

    // mask + online softmax (log2 domain)
    const int kbase = j * kBN;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < kBN / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kbase + nb * 8 + 2 * t + (e & 1);
        const int qr = qrow0 + (e >> 1) * 8;
        const bool ok = key < L && (!causal || key <= qr);
        const float x = ok ? sacc[nb][e] * scale_log2 : -INFINITY;
        sacc[nb][e] = x;
        mx[e >> 1] = fmaxf(mx[e >> 1], x);
      }
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_run[r], mx[r]);
      msafe[r] = mnew == -INFINITY ? 0.f : mnew;
      corr[r] = exp2f(m_run[r] - msafe[r]);  // m_run = -inf -> 0
      m_run[r] = mnew;
      l_run[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < DBLK; ++i) {
      oacc[i][0] *= corr[0];
      oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1];
      oacc[i][3] *= corr[1];
    }
    uint32_t pf[kBN / 16][4];
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nb = 0; nb < kBN / 8; ++nb) {
      const float p0 = exp2f(sacc[nb][0] - msafe[0]);
      const float p1 = exp2f(sacc[nb][1] - msafe[0]);
      const float p2 = exp2f(sacc[nb][2] - msafe[1]);
      const float p3 = exp2f(sacc[nb][3] - msafe[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      // C-fragment of two adjacent 8-key blocks == A-fragment of one 16-key step
      pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
    l_run[0] += rs[0];
    l_run[1] += rs[1];

    // O += P V
    const uint32_t sVs = sV + st * kBN * ROWB;
#pragma unroll
    for (int kk = 0; kk < kBN / 16; ++kk) {
#pragma unroll
      for (int db = 0; db < DBLK / 2; ++db) {
        uint32_t b0, b1, b2, b3;
        const int row = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int chunk = db * 2 + (lane >> 4);
        ldmatrix_x4_trans(sVs + swz<D>(row, chunk), b0, b1, b2, b3);
        mma_bf16(oacc[2 * db], pf[kk], b0, b1);
        mma_bf16(oacc[2 * db + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // all warps done with stage `st` before it is refilled
  }

  // finalize: row sums are spread over the 4 threads of a quad
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qr = qrow0 + r * 8;
    if (qr < L) {
      const float inv = l_run[r] > 0.f ? 1.0f / l_run[r] : 0.f;
      __nv_bfloat16* orow = o + static_cast<long long>(s0 + qr) * ldo + h * D;
#pragma unroll
      for (int db = 0; db < DBLK; ++db) {
        *reinterpret_cast<uint32_t*>(orow + db * 8 + 2 * t) =
            pack_bf16(oacc[db][2 * r] * inv, oacc[db][2 * r + 1] * inv);
      }
    }
  }
}

template <int D>
int launch_attn(const void* q, const void* k, const void* v, long long ld, void* o, long long ldo, const int* cu,
                int num_seqs, int max_seqlen, int Hq, int Hkv, float scale, int causal, cudaStream_t stream) {
  constexpr int smem = (kBM + 4 * kBN) * D * 2;
  auto kern = attn_fwd_kernel<D>;
  static bool attr_set = false;
  if (!attr_set) {
    B2S_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid((max_seqlen + kBM - 1) / kBM, Hq, num_seqs);
  kern<<<grid, kAttnThreads, smem, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(q), reinterpret_cast<const __nv_bfloat16*>(k),
      reinterpret_cast<const __nv_bfloat16*>(v), ld, reinterpret_cast<__nv_bfloat16*>(o), ldo, cu, Hq, Hkv,
      scale * 1.4426950408889634f, causal);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace

// attention_tc.cu
int attention_fwd_tc(const void* q, const void* k, const void* v, long long ld_qkv, void* o, long long ld_o,
                     const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv, int D,
                     float scale, int causal, float* lse, cudaStream_t stream, const AttnDrop* drop);

namespace {
int g_attn_impl = 1;  // 1 = tcgen05 / TMEM / TMA kernel (attention_tc.cu), 0 = legacy mma.sync kernel (this file)
}
void attention_set_impl(int impl) { g_attn_impl = impl; }
int attention_get_impl() { return g_attn_impl; }

int attention_fwd(const void* q, const void* k, const void* v, long long ld_qkv, void* o, long long ld_o,
                  const int* cu_seqlens, int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv, int D,
                  float scale, int causal, float* lse, cudaStream_t stream, const AttnDrop* drop) {
  B2S_REQUIRE(q && k && v && o && cu_seqlens, "attention_fwd: null pointer");
  if (drop != nullptr && drop->thresh == 0u) drop = nullptr;
  B2S_REQUIRE(total_rows > 0, "attention_fwd: total_rows must be the row count of the packed q/k/v buffers");
  B2S_REQUIRE(num_seqs > 0 && max_seqlen > 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "attention_fwd: bad head counts");
  B2S_REQUIRE(ld_qkv % 8 == 0 && ld_o % 2 == 0, "attention_fwd: strides must keep 16-byte row alignment");
  B2S_REQUIRE((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(v) & 15) == 0,
              "attention_fwd: q/k/v must be 16-byte aligned");
  if (g_attn_impl == 1 && (D == 64 || D == 128) && (ld_qkv * 2) % 16 == 0 && ld_o % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
    return attention_fwd_tc(q, k, v, ld_qkv, o, ld_o, cu_seqlens, num_seqs, max_seqlen, total_rows, Hq, Hkv, D, scale,
                            causal, lse, stream, drop);
  }
  B2S_REQUIRE(drop == nullptr, "attention_fwd: attention dropout needs the tcgen05 kernel (impl 1, D = 64 / 128)");
  B2S_REQUIRE(lse == nullptr, "attention_fwd: the log-sum-exp output needs the tcgen05 kernel (impl 1, D = 64 / 128)");
  if (D == 64) return launch_attn<64>(q, k, v, ld_qkv, o, ld_o, cu_seqlens, num_seqs, max_seqlen, Hq, Hkv, scale, causal, stream);
  if (D == 128) return launch_attn<128>(q, k, v, ld_qkv, o, ld_o, cu_seqlens, num_seqs, max_seqlen, Hq, Hkv, scale, causal, stream);
  set_last_error("attention_fwd: head_dim %d unsupported (64 or 128)", D);
  return B2S_ERR_UNSUPPORTED;
}

}  // namespace b2s
