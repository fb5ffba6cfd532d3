// misc.cu -- the remaining memory-bound pieces of the path:
//   * conv0_ln_gelu: HuBERT feature-extractor layer 0, Conv1d(1->512,k10,s5)+bias -> LayerNorm(512) -> erf-GELU,
//     all in registers, channels-last bf16 out (TF/models/hubert/modeling_hubert.py:127-151, layer_id 0).
//   * embed_splice: builds the packed LLM input sequence (prefix | audio or text | suffix[1:] | response[1:])
//     by gathering embed_tokens rows and the projected audio rows (REF/utils.py:27-46,85-164).
//   * rowpair_sqdiff: per-row sum of squared differences for the feature-distillation MSE (REF/trainer.py:358-370).
//   * posconv_weight_pack: weight-norm (dim=2) of the positional conv + repack to the K-major layout the
//     tap-walk GEMM consumes (TF/models/hubert/modeling_hubert.py:45-92).
#include <climits>

#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kConv0Out = 512;
constexpr int kConv0K = 10;
constexpr int kConv0S = 5;
constexpr int kConv0TT = 4;  // time steps per warp iteration

__global__ void __launch_bounds__(256)
conv0_ln_gelu_kernel(const float* __restrict__ wave, long long wave_stride, int samples, const float* __restrict__ w,
                     const float* __restrict__ bias, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, __nv_bfloat16* __restrict__ y, int out_frames, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  __shared__ float ws[kConv0K][kConv0Out];  // transposed taps: ws[j][c]
  for (int i = threadIdx.x; i < kConv0K * kConv0Out; i += blockDim.x) {
    const int c = i / kConv0K, j = i - c * kConv0K;
    ws[j][c] = w[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y;
  const int t0 = (blockIdx.x * 8 + warp) * kConv0TT;
  if (t0 >= out_frames) return;
  const float* x = wave + static_cast<long long>(b) * wave_stride + static_cast<long long>(t0) * kConv0S;

  // 25 input samples cover 4 consecutive windows
  float xs[kConv0S * (kConv0TT - 1) + kConv0K];
#pragma unroll
  for (int i = 0; i < kConv0S * (kConv0TT - 1) + kConv0K; ++i) {
    const long long idx = static_cast<long long>(t0) * kConv0S + i;
    xs[i] = idx < samples ? __ldg(x + i) : 0.f;
  }
  float acc[kConv0TT][16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 bb = *reinterpret_cast<const float2*>(bias + 64 * i + 2 * lane);
#pragma unroll
    for (int tt = 0; tt < kConv0TT; ++tt) {
      acc[tt][2 * i] = bb.x;
      acc[tt][2 * i + 1] = bb.y;
    }
  }
#pragma unroll
  for (int j = 0; j < kConv0K; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 ww = *reinterpret_cast<const float2*>(&ws[j][64 * i + 2 * lane]);
#pragma unroll
      for (int tt = 0; tt < kConv0TT; ++tt) {
        acc[tt][2 * i] = fmaf(ww.x, xs[tt * kConv0S + j], acc[tt][2 * i]);
        acc[tt][2 * i + 1] = fmaf(ww.y, xs[tt * kConv0S + j], acc[tt][2 * i + 1]);
      }
    }
  }
#pragma unroll
  for (int tt = 0; tt < kConv0TT; ++tt) {
    const int t = t0 + tt;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[tt][i];
    const float mean = warp_sum(s) * (1.0f / kConv0Out);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d = acc[tt][i] - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / kConv0Out) + eps);
    if (t < out_frames) {
      __nv_bfloat16* yo = y + (static_cast<long long>(b) * out_frames + t) * kConv0Out;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 gm = *reinterpret_cast<const float2*>(gamma + 64 * i + 2 * lane);
        const float2 bt = *reinterpret_cast<const float2*>(beta + 64 * i + 2 * lane);
        const float o0 = gelu_erf((acc[tt][2 * i] - mean) * rstd * gm.x + bt.x);
        const float o1 = gelu_erf((acc[tt][2 * i + 1] - mean) * rstd * gm.y + bt.y);
        *reinterpret_cast<uint32_t*>(yo + 64 * i + 2 * lane) = pack_h16(o0, o1, f16);
      }
    }
  }
}

// one warp per output row; C % 256 == 0
__global__ void __launch_bounds__(256)
embed_splice_kernel(const __nv_bfloat16* __restrict__ table, const float* __restrict__ audio,
                    const int* __restrict__ row_src, float* __restrict__ h0, long long rows, int C, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int src = row_src[row];
  float* dst = h0 + row * C;
  if (src == INT_MIN) {  // left-padding row of the reference's batched layout (REF/utils.py:136-146)
    for (int i = lane; i < C / 4; i += 32) reinterpret_cast<float4*>(dst)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (src >= 0) {
    const uint4* s = reinterpret_cast<const uint4*>(table + static_cast<long long>(src) * C);
    for (int i = lane; i < C / 8; i += 32) {
      float f[8];
      unpack8_h16(__ldg(s + i), f, f16);
      reinterpret_cast<float4*>(dst)[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
      reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
    }
  } else {
    const float4* s = reinterpret_cast<const float4*>(audio + static_cast<long long>(-(src + 1)) * C);
    for (int i = lane; i < C / 4; i += 32) reinterpret_cast<float4*>(dst)[i] = __ldg(s + i);
  }
}

__global__ void __launch_bounds__(256)
rowpair_sqdiff_kernel(const float* __restrict__ h, const int* __restrict__ rows_a, const int* __restrict__ rows_b,
                      float* __restrict__ out, int pairs, int C) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= pairs) return;
  const float4* a = reinterpret_cast<const float4*>(h + static_cast<long long>(rows_a[i]) * C);
  const float4* b = reinterpret_cast<const float4*>(h + static_cast<long long>(rows_b[i]) * C);
  float s = 0.f;
  for (int k = lane; k < C / 4; k += 32) {
    const float4 x = a[k], y = b[k];
    const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
    s += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  s = warp_sum(s);
  if (lane == 0) out[i] = s;
}

// one CTA per tap k: norm over (co, ci) of v[:, :, k], then scaled bf16 write to [co][k][ci]
__global__ void __launch_bounds__(256)
posconv_weight_pack_kernel(const float* __restrict__ g, const float* __restrict__ v, uint16_t* __restrict__ wp,
                           int cout, int cin_g, int K, int f16) {
  const int k = blockIdx.x;
  const long long n = static_cast<long long>(cout) * cin_g;
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i * K + k];
    s = fmaf(x, x, s);
  }
  __shared__ float sh[8];
  __shared__ float scale_sh;
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    scale_sh = g[k] / sqrtf(t);
  }
  __syncthreads();
  const float scale = scale_sh;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const long long co = i / cin_g, ci = i - co * cin_g;
    wp[(co * K + k) * cin_g + ci] = float_to_h16(v[i * K + k] * scale, f16);
  }
}

// log-mel (B, C, T) fp32 -> channels-last bf16 (B, T + 2, C) with a zero row before and after each utterance
// (the zero padding of Whisper's conv1/conv2, TF/models/whisper/modeling_whisper.py:566-567)
__global__ void mel_to_padded_cl_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, int B, int C, int T,
                                        int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * (T + 2) * C;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const long long r = i / C;
  const int tp = static_cast<int>(r % (T + 2));
  const int b = static_cast<int>(r / (T + 2));
  float v = 0.f;
  if (tp >= 1 && tp <= T) v = x[(static_cast<long long>(b) * C + c) * T + (tp - 1)];
  y[i] = float_to_h16(v, f16);
}

__global__ void cast_f32_h16_kernel(const float* __restrict__ x, uint16_t* __restrict__ y, long long n, int f16) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(x + i);
    uint2 u;
    u.x = pack_h16(v.x, v.y, f16);
    u.y = pack_h16(v.z, v.w, f16);
    *reinterpret_cast<uint2*>(y + i) = u;
  } else {
    for (long long j = i; j < n; ++j) y[j] = float_to_h16(x[j], f16);
  }
}

__global__ void cast_h16_f32_kernel(const uint16_t* __restrict__ x, float* __restrict__ y, long long n, int f16) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) y[i] = h16_to_float(x[i], f16);
}

}  // namespace

int conv0_ln_gelu_fwd(const float* wave, long long wave_stride, int batches, int samples, const float* w,
                      const float* bias, const float* gamma, const float* beta, float eps, void* y_bf16, int out_frames,
                      int fmt, cudaStream_t stream) {
  B2S_REQUIRE(wave && w && bias && gamma && beta && y_bf16, "conv0_ln_gelu_fwd: null pointer");
  B2S_REQUIRE(batches > 0 && samples >= kConv0K, "conv0_ln_gelu_fwd: need at least %d samples", kConv0K);
  B2S_REQUIRE(out_frames == (samples - kConv0K) / kConv0S + 1, "conv0_ln_gelu_fwd: out_frames mismatch");
  dim3 grid((out_frames + 8 * kConv0TT - 1) / (8 * kConv0TT), batches);
  conv0_ln_gelu_kernel<<<grid, 256, 0, stream>>>(wave, wave_stride, samples, w, bias, gamma, beta, eps,
                                                 reinterpret_cast<__nv_bfloat16*>(y_bf16), out_frames, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int embed_splice_fwd(const void* embed_table_bf16, const float* audio_embeds, const int* row_src, float* h0,
                     long long rows, int C, int fmt, cudaStream_t stream) {
  B2S_REQUIRE(embed_table_bf16 && row_src && h0, "embed_splice_fwd: null pointer");
  B2S_REQUIRE(C > 0 && C % 8 == 0, "embed_splice_fwd: C must be a multiple of 8");
  if (rows <= 0) return B2S_OK;
  embed_splice_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(embed_table_bf16), audio_embeds, row_src, h0, rows, C, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int rowpair_sqdiff_fwd(const float* h, const int* rows_a, const int* rows_b, float* out, int pairs, int C,
                       cudaStream_t stream) {
  B2S_REQUIRE(h && rows_a && rows_b && out, "rowpair_sqdiff_fwd: null pointer");
  B2S_REQUIRE(C > 0 && C % 4 == 0, "rowpair_sqdiff_fwd: C must be a multiple of 4");
  if (pairs <= 0) return B2S_OK;
  rowpair_sqdiff_kernel<<<(pairs + 7) / 8, 256, 0, stream>>>(h, rows_a, rows_b, out, pairs, C);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int posconv_weight_pack(const float* g, const float* v, void* w_packed_bf16, int cout, int cin_g, int k, int fmt,
                        cudaStream_t stream) {
  B2S_REQUIRE(g && v && w_packed_bf16 && cout > 0 && cin_g > 0 && k > 0, "posconv_weight_pack: bad arguments");
  posconv_weight_pack_kernel<<<k, 256, 0, stream>>>(g, v, reinterpret_cast<uint16_t*>(w_packed_bf16), cout, cin_g, k,
                                                    fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int mel_to_padded_cl(const float* x, void* y_bf16, int batches, int channels, int frames, int fmt,
                     cudaStream_t stream) {
  B2S_REQUIRE(x && y_bf16 && batches > 0 && channels > 0 && frames > 0, "mel_to_padded_cl: bad arguments");
  const long long total = static_cast<long long>(batches) * (frames + 2) * channels;
  mel_to_padded_cl_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(
      x, reinterpret_cast<uint16_t*>(y_bf16), batches, channels, frames, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

// out[i, :] = src[index[i], :] for bf16 rows (C % 8 == 0): warp per row, 16-byte accesses
__global__ void __launch_bounds__(256)
gather_rows_bf16_kernel(const __nv_bfloat16* __restrict__ src, const int* __restrict__ index,
                        __nv_bfloat16* __restrict__ out, long long rows, int C) {
  pdl_trigger();  // PDL: let a dependent GEMM take the SMs this grid frees (b2s_common.cuh)
  const int lane = threadIdx.x & 31;
  const long long i = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (i >= rows) return;
  const long long s = index[i];
  for (int k = lane * 8; k < C; k += 256)
    *reinterpret_cast<uint4*>(out + i * C + k) = *reinterpret_cast<const uint4*>(src + s * C + k);
}

int gather_rows_bf16(const void* src, const int* index, void* out, long long rows, int C, cudaStream_t stream) {
  B2S_REQUIRE(src && index && out, "gather_rows_bf16: null pointer");
  B2S_REQUIRE(C % 8 == 0, "gather_rows_bf16: C must be a multiple of 8");
  if (rows <= 0) return B2S_OK;
  gather_rows_bf16_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), index, reinterpret_cast<__nv_bfloat16*>(out), rows, C);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int cast_f32_to_h16(const float* x, void* y, long long n, int fmt, cudaStream_t stream) {
  if (n <= 0) return B2S_OK;
  B2S_REQUIRE(x && y, "cast_f32_to_h16: null pointer");
  const long long threads = (n + 3) / 4;
  cast_f32_h16_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(
      x, reinterpret_cast<uint16_t*>(y), n, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int cast_h16_to_f32(const void* x, float* y, long long n, int fmt, cudaStream_t stream) {
  if (n <= 0) return B2S_OK;
  B2S_REQUIRE(x && y, "cast_h16_to_f32: null pointer");
  cast_h16_f32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint16_t*>(x), y, n, fmt);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
