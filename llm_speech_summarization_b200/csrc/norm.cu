// norm.cu -- LayerNorm / RMSNorm family (HBM-bound): one warp per row, the row held in registers
// (two-pass mean/variance in fp32), 16/32-byte vector loads, bf16 vector stores, warp-shuffle reductions.
//
// Replaces: the nn.LayerNorm calls inside HubertLayerNormConvLayer / HubertFeatureProjection /
// HubertEncoderLayerStableLayerNorm (TF/models/hubert/modeling_hubert.py:127-151,216-231,505-548,613),
// LlamaRMSNorm (TF/models/llama/modeling_llama.py:53-67) and the AvgPool1d of
// REF/model/audio_encoder.py:59-63 (fused with the encoder's final LayerNorm).
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kWarpsPerCta = 8;

template <bool IN_BF16>
__device__ __forceinline__ void load8(const void* base, long long elem_off, float (&f)[8], int f16 = 0) {
  if constexpr (IN_BF16) {  // 16-bit input in format f16 (0 = bf16, 1 = fp16)
    ld8h(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off, f, f16);
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + elem_off);
    const float4 a = p[0], b = p[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
}

__device__ __forceinline__ void store8_h16(void* base, long long elem_off, const float (&f)[8], int f16) {
  st8h(reinterpret_cast<__nv_bfloat16*>(base) + elem_off, f, f16);
}

// row -> registers, returns mean and rstd (two-pass)
template <int GROUPS, bool IN_BF16>
__device__ __forceinline__ void load_row_stats(const void* x, long long row_off, int lane, float (&v)[GROUPS][8],
                                               float eps, float& mean, float& rstd, int f16 = 0) {
  constexpr int C = GROUPS * 256;
  float s = 0.f;
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    load8<IN_BF16>(x, row_off + (g * 32 + lane) * 8, v[g], f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[g][j];
  }
  mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[g][j] - mean;
      q = fmaf(d, d, q);
    }
  }
  rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
}

template <int GROUPS, bool IN_BF16, bool GELU>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_kernel(const void* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 float eps, void* __restrict__ y, long long rows, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  pdl_wait();     // ... and this grid itself may have been launched early: the predecessor's rows are complete from here
  constexpr int C = GROUPS * 256;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[GROUPS][8];
  float mean, rstd;
  load_row_stats<GROUPS, IN_BF16>(x, row * C, lane, v, eps, mean, rstd, f16);
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    float gm[8], bt[8], o[8];
    load8<false>(gamma, (g * 32 + lane) * 8, gm);
    load8<false>(beta, (g * 32 + lane) * 8, bt);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = (v[g][j] - mean) * rstd * gm[j] + bt[j];
      o[j] = GELU ? gelu_erf(t) : t;
    }
    store8_h16(y, row * C + (g * 32 + lane) * 8, o, f16);
  }
}

template <int GROUPS, bool GATHER>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
rmsnorm_kernel(const float* __restrict__ x, const int* __restrict__ row_index, const float* __restrict__ w, float eps,
               void* __restrict__ y, long long rows, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  pdl_wait();     // ... and this grid itself may have been launched early: the predecessor's rows are complete from here
  constexpr int C = GROUPS * 256;
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;
  const long long src = GATHER ? static_cast<long long>(row_index[row]) : row;
  float v[GROUPS][8];
  float q = 0.f;
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    load8<false>(x, src * C + (g * 32 + lane) * 8, v[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) q = fmaf(v[g][j], v[g][j], q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    float wt[8], o[8];
    load8<false>(w, (g * 32 + lane) * 8, wt);
    // LlamaRMSNorm: weight * (x * rstd) with the normalised value rounded to the activation dtype first
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = wt[j] * (v[g][j] * rstd);
    store8_h16(y, row * C + (g * 32 + lane) * 8, o, f16);
  }
}

template <int GROUPS>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
layernorm_avgpool_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                         float eps, void* __restrict__ y, int batches, int frames, int kernel, int stride,
                         int out_frames, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  pdl_wait();     // ... and this grid itself may have been launched early: the predecessor's rows are complete from here
  constexpr int C = GROUPS * 256;
  const int lane = threadIdx.x & 31;
  const long long orow = static_cast<long long>(blockIdx.x) * kWarpsPerCta + (threadIdx.x >> 5);
  if (orow >= static_cast<long long>(batches) * out_frames) return;
  const int b = static_cast<int>(orow / out_frames);
  const int j = static_cast<int>(orow - static_cast<long long>(b) * out_frames);
  float acc[GROUPS][8];
#pragma unroll
  for (int g = 0; g < GROUPS; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
  for (int r = 0; r < kernel; ++r) {
    const long long irow = static_cast<long long>(b) * frames + j * stride + r;
    float v[GROUPS][8];
    float mean, rstd;
    load_row_stats<GROUPS, false>(x, irow * C, lane, v, eps, mean, rstd);
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[g][i] += (v[g][i] - mean) * rstd;
  }
  const float inv = 1.0f / kernel;
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    float gm[8], bt[8], o[8];
    load8<false>(gamma, (g * 32 + lane) * 8, gm);
    load8<false>(beta, (g * 32 + lane) * 8, bt);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = acc[g][i] * inv * gm[i] + bt[i];  // mean of affine = affine of mean
    store8_h16(y, orow * C + (g * 32 + lane) * 8, o, f16);
  }
}

#define B2S_DISPATCH_GROUPS(C, ...)                                               \
  switch ((C) / 256) {                                                            \
    case 1: { constexpr int G = 1; __VA_ARGS__; break; }                          \
    case 2: { constexpr int G = 2; __VA_ARGS__; break; }                          \
    case 3: { constexpr int G = 3; __VA_ARGS__; break; }                          \
    case 4: { constexpr int G = 4; __VA_ARGS__; break; }                          \
    case 5: { constexpr int G = 5; __VA_ARGS__; break; }                          \
    case 8: { constexpr int G = 8; __VA_ARGS__; break; }                          \
    case 12: { constexpr int G = 12; __VA_ARGS__; break; }                        \
    case 16: { constexpr int G = 16; __VA_ARGS__; break; }                        \
    default:                                                                      \
      set_last_error("norm: unsupported width %d (multiples of 256: 256..1280, 2048, 3072, 4096)", (C)); \
      return B2S_ERR_UNSUPPORTED;                                                 \
  }

}  // namespace

int layernorm_fwd(const void* x, int in_bf16, const float* gamma, const float* beta, float eps, int act_gelu,
                  void* y_bf16, long long rows, int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x && gamma && beta && y_bf16, "layernorm_fwd: null pointer");
  B2S_REQUIRE(C > 0 && C % 256 == 0, "layernorm_fwd: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  const unsigned grid = static_cast<unsigned>((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  int rc_ = B2S_OK;
  B2S_DISPATCH_GROUPS(C, {
    if (in_bf16) {
      if (act_gelu) rc_ = launch_pdl_kernel(layernorm_kernel<G, true, true>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x, gamma, beta, eps, y_bf16, rows, fmt);
      else rc_ = launch_pdl_kernel(layernorm_kernel<G, true, false>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x, gamma, beta, eps, y_bf16, rows, fmt);
    } else {
      if (act_gelu) rc_ = launch_pdl_kernel(layernorm_kernel<G, false, true>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x, gamma, beta, eps, y_bf16, rows, fmt);
      else rc_ = launch_pdl_kernel(layernorm_kernel<G, false, false>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream, x, gamma, beta, eps, y_bf16, rows, fmt);
    }
  });
  if (rc_ != B2S_OK) return rc_;
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int rmsnorm_fwd(const float* x, const float* w, float eps, void* y_bf16, long long rows, int C, int fmt,
                cudaStream_t stream) {
  B2S_REQUIRE(x && w && y_bf16, "rmsnorm_fwd: null pointer");
  B2S_REQUIRE(C > 0 && C % 256 == 0, "rmsnorm_fwd: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  const unsigned grid = static_cast<unsigned>((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  int rc_ = B2S_OK;
  const int* no_index = nullptr;
  B2S_DISPATCH_GROUPS(C, (rc_ = launch_pdl_kernel(rmsnorm_kernel<G, false>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream,
                                                  x, no_index, w, eps, y_bf16, rows, fmt)));
  if (rc_ != B2S_OK) return rc_;
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int rmsnorm_gather_fwd(const float* x, const int* row_index, const float* w, float eps, void* y_bf16, long long rows,
                       int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x && w && y_bf16 && row_index, "rmsnorm_gather_fwd: null pointer");
  B2S_REQUIRE(C > 0 && C % 256 == 0, "rmsnorm_gather_fwd: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  const unsigned grid = static_cast<unsigned>((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  int rc_ = B2S_OK;
  B2S_DISPATCH_GROUPS(C, (rc_ = launch_pdl_kernel(rmsnorm_kernel<G, true>, dim3(grid), dim3(kWarpsPerCta * 32), 0, stream,
                                                  x, row_index, w, eps, y_bf16, rows, fmt)));
  if (rc_ != B2S_OK) return rc_;
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int layernorm_avgpool_fwd(const float* x, const float* gamma, const float* beta, float eps, void* y_bf16, int batches,
                          int frames, int C, int kernel, int stride, int out_frames, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(x && gamma && beta && y_bf16, "layernorm_avgpool_fwd: null pointer");
  B2S_REQUIRE(C > 0 && C % 256 == 0 && kernel > 0 && stride > 0, "layernorm_avgpool_fwd: bad sizes");
  B2S_REQUIRE(out_frames >= 0 && (out_frames == 0 || (out_frames - 1) * stride + kernel <= frames),
              "layernorm_avgpool_fwd: pooling window exceeds the input");
  const long long orows = static_cast<long long>(batches) * out_frames;
  if (orows <= 0) return B2S_OK;
  const unsigned grid = static_cast<unsigned>((orows + kWarpsPerCta - 1) / kWarpsPerCta);
  B2S_DISPATCH_GROUPS(C, (layernorm_avgpool_kernel<G><<<grid, kWarpsPerCta * 32, 0, stream>>>(
                             x, gamma, beta, eps, y_bf16, batches, frames, kernel, stride, out_frames, fmt)));
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
