// regularize.cu -- memory-bound pieces of the train-mode regularisers of the HuBERT encoder (SURVEY.md section 8 rows
// a6 / f4): the dropout sites that are not GEMM epilogues, SpecAugment's time masking and their backward.
//   TF/models/hubert/modeling_hubert.py:585-587 (dropout after hidden + positional conv), :842-886 (_mask_hidden_states),
//   :223-230 (feature-projection dropout). The keep/drop decisions are regenerated from counters (rng.cuh), never stored.
#include "b2s_common.cuh"
#include "ops.cuh"
#include "rng.cuh"

namespace b2s {
namespace {

// x (fp32 and / or bf16 copy of the same logical tensor) *= keep ? 1/(1-p) : 0, element index = linear index
__global__ void __launch_bounds__(256)
dropout_apply_kernel(float* __restrict__ xf, __nv_bfloat16* __restrict__ xb, long long n8, DropSpec d, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint32_t e0 = static_cast<uint32_t>(i * 8);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = rng_keep(e0 + j, d.k1, d.k2, d.thresh) ? d.inv_keep : 0.f;
  if (xf != nullptr) {
    float v[8];
    ld8f(xf + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= m[j];
    st8f(xf + i * 8, v);
  }
  if (xb != nullptr) {
    float v[8];
    ld8h(xb + i * 8, v, f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= m[j];
    st8h(xb + i * 8, v, f16);
  }
}

// SpecAugment along time: h[row, :] = embed[:] where time_mask[row]   (warp per row)
__global__ void __launch_bounds__(256)
mask_rows_kernel(float* __restrict__ h, const unsigned char* __restrict__ time_mask, const float* __restrict__ embed,
                 long long rows, int C) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows || time_mask[row] == 0) return;
  for (int c = lane * 4; c < C; c += 128)
    *reinterpret_cast<float4*>(h + row * C + c) = *reinterpret_cast<const float4*>(embed + c);
}

// backward of [feature-projection dropout -> SpecAugment]: rows replaced by masked_spec_embed send their gradient to
// the embedding and nothing upstream; the others get the dropout mask of the feature-projection site
__global__ void __launch_bounds__(256)
featproj_reg_bwd_kernel(float* __restrict__ dh, const unsigned char* __restrict__ time_mask, float* __restrict__ g_embed,
                        long long rows, int C, DropSpec d) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const bool masked = time_mask != nullptr && time_mask[row] != 0;
  if (!masked && d.thresh == 0u) return;
  for (int c = lane * 4; c < C; c += 128) {
    float4* p = reinterpret_cast<float4*>(dh + row * C + c);
    float4 v = *p;
    if (masked) {
      if (g_embed != nullptr) {
        atomicAdd(g_embed + c, v.x);
        atomicAdd(g_embed + c + 1, v.y);
        atomicAdd(g_embed + c + 2, v.z);
        atomicAdd(g_embed + c + 3, v.w);
      }
      v = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const uint32_t e0 = static_cast<uint32_t>(row * C + c);
      v.x = rng_keep(e0, d.k1, d.k2, d.thresh) ? v.x * d.inv_keep : 0.f;
      v.y = rng_keep(e0 + 1, d.k1, d.k2, d.thresh) ? v.y * d.inv_keep : 0.f;
      v.z = rng_keep(e0 + 2, d.k1, d.k2, d.thresh) ? v.z * d.inv_keep : 0.f;
      v.w = rng_keep(e0 + 3, d.k1, d.k2, d.thresh) ? v.w * d.inv_keep : 0.f;
    }
    *p = v;
  }
}

// keep-mask dump for the parity tests: out[i] = 1 if element i of the stream is kept
__global__ void __launch_bounds__(256)
drop_mask_dump_kernel(unsigned char* __restrict__ out, long long n, uint32_t e_first, DropSpec d) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rng_keep(e_first + static_cast<uint32_t>(i), d.k1, d.k2, d.thresh) ? 1 : 0;
}

}  // namespace

int dropout_apply(float* x_f32, void* x_bf16, long long n, const DropSpec& d, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x_f32 || x_bf16, "dropout_apply: null pointer");
  B2S_REQUIRE(n % 8 == 0 && n < (1LL << 32), "dropout_apply: element count must be a multiple of 8 below 2^32");
  if (n <= 0 || d.thresh == 0u) return B2S_OK;
  dropout_apply_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(
      x_f32, reinterpret_cast<__nv_bfloat16*>(x_bf16), n / 8, d, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int mask_rows_f32(float* h, const unsigned char* time_mask, const float* embed, long long rows, int C,
                  cudaStream_t stream) {
  B2S_REQUIRE(h && time_mask && embed, "mask_rows_f32: null pointer");
  B2S_REQUIRE(C % 4 == 0, "mask_rows_f32: C must be a multiple of 4");
  if (rows <= 0) return B2S_OK;
  mask_rows_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(h, time_mask, embed, rows, C);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int featproj_reg_bwd(float* dh, const unsigned char* time_mask, float* g_embed, long long rows, int C, const DropSpec& d,
                     cudaStream_t stream) {
  B2S_REQUIRE(dh, "featproj_reg_bwd: null pointer");
  B2S_REQUIRE(C % 4 == 0 && rows * C < (1LL << 32), "featproj_reg_bwd: C %% 4 and rows*C < 2^32 required");
  if (rows <= 0 || (time_mask == nullptr && d.thresh == 0u)) return B2S_OK;
  featproj_reg_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(dh, time_mask, g_embed, rows, C, d);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int drop_mask_dump(unsigned char* out, long long n, unsigned long long seed, uint32_t site, uint32_t a, uint32_t b,
                   float p, uint32_t e_first, cudaStream_t stream) {
  B2S_REQUIRE(out, "drop_mask_dump: null pointer");
  if (n <= 0) return B2S_OK;
  DropSpec d{};
  d.thresh = drop_threshold(p);
  d.inv_keep = 1.f;
  rng_stream_key(seed, site, a, b, &d.k1, &d.k2);
  drop_mask_dump_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(out, n, e_first, d);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
