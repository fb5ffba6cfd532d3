// backward_enc.cu -- memory-bound backward kernels of the trainable audio encoder (REF/trainer.py:98-105 puts every
// AudioEncoder parameter in the optimizer; autograd through TF/models/hubert/modeling_hubert.py:127-231,505-624 and
// REF/model/audio_encoder.py:56-88). The dense contractions (dgrad / wgrad) run on the tcgen05 GEMM with MN-major
// operands; what is left is here:
//   layernorm_bwd_ex : LayerNorm (+ optional erf-GELU on its output) backward with parameter gradients accumulated
//                      in registers across rows (one atomic per column per block)
//   colsum_accum     : bias gradients
//   avgpool_bwd      : AvgPool1d over time, transposed
//   col2im_add       : strided Conv1d data gradient = gather of the per-tap dgrad GEMM columns
//   conv0_bwd        : conv layer 0 (1 -> 512 channels, CUDA cores) fused with its LayerNorm + GELU backward
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kWarps = 8;

template <int GROUPS, bool GELU>
__global__ void __launch_bounds__(kWarps * 32)
layernorm_bwd_ex_kernel(const void* __restrict__ x, int x_bf16, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float eps, const void* __restrict__ dy, int dy_bf16, float* dh,
                        int accumulate, __nv_bfloat16* dx_bf16, float* __restrict__ dgamma, float* __restrict__ dbeta,
                        long long rows) {
  constexpr int C = GROUPS * 256;
  __shared__ float s_red[kWarps][C];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gm[GROUPS][8], bt[GROUPS][8], adg[GROUPS][8], adb[GROUPS][8];
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {
    const int c = (g * 32 + lane) * 8;
    ld8f(gamma + c, gm[g]);
    if constexpr (GELU) ld8f(beta + c, bt[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) adg[g][j] = adb[g][j] = 0.f;
  }
  for (long long row = static_cast<long long>(blockIdx.x) * kWarps + warp; row < rows;
       row += static_cast<long long>(gridDim.x) * kWarps) {
    float xv[GROUPS][8], dv[GROUPS][8];
    float s = 0.f;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      if (x_bf16) ld8bf(reinterpret_cast<const __nv_bfloat16*>(x) + row * C + c, xv[g]);
      else ld8f(reinterpret_cast<const float*>(x) + row * C + c, xv[g]);
      if (dy_bf16) ld8bf(reinterpret_cast<const __nv_bfloat16*>(dy) + row * C + c, dv[g]);
      else ld8f(reinterpret_cast<const float*>(dy) + row * C + c, dv[g]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += xv[g][j];
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xv[g][j] -= mean;
        q = fmaf(xv[g][j], xv[g][j], q);
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = xv[g][j] * rstd;
        float d = dv[g][j];
        if constexpr (GELU) d *= gelu_erf_grad(fmaf(gm[g][j], xh, bt[g][j]));
        adg[g][j] = fmaf(d, xh, adg[g][j]);
        adb[g][j] += d;
        d *= gm[g][j];
        xv[g][j] = xh;
        dv[g][j] = d;
        sg += d;
        sgx = fmaf(d, xh, sgx);
      }
    sg = warp_sum(sg) * (1.0f / C);
    sgx = warp_sum(sgx) * (1.0f / C);
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const int c = (g * 32 + lane) * 8;
      float acc[8];
      if (dh != nullptr && accumulate) ld8f(dh + row * C + c, acc);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += rstd * (dv[g][j] - sg - xv[g][j] * sgx);
      if (dh != nullptr) st8f(dh + row * C + c, acc);
      if (dx_bf16 != nullptr) st8bf(dx_bf16 + row * C + c, acc);
    }
  }
  // parameter gradients: registers -> shared (per warp) -> one atomic per column per block
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) st8f(&s_red[warp][(g * 32 + lane) * 8], pass == 0 ? adg[g] : adb[g]);
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += s_red[w][i];
      atomicAdd((pass == 0 ? dgamma : dbeta) + i, t);
    }
    __syncthreads();
  }
}

// out[c] += sum_r x[r, c]
__global__ void __launch_bounds__(kWarps * 32)
colsum_kernel(const void* __restrict__ x, int x_bf16, float* __restrict__ out, long long rows, int C) {
  __shared__ float s_red[kWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long row = static_cast<long long>(blockIdx.y) * kWarps + warp; row < rows;
       row += static_cast<long long>(gridDim.y) * kWarps) {
    float v[8];
    if (x_bf16) ld8bf(reinterpret_cast<const __nv_bfloat16*>(x) + row * C + c, v);
    else ld8f(reinterpret_cast<const float*>(x) + row * C + c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
  st8f(&s_red[warp][lane * 8], acc);
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += s_red[w][threadIdx.x];
  atomicAdd(out + blockIdx.x * 256 + threadIdx.x, t);
}

// dx[b, t, :] = (1 / kernel) * sum_{j : j*stride <= t < j*stride + kernel} dpooled[b, j, :]
__global__ void __launch_bounds__(256)
avgpool_bwd_kernel(const float* __restrict__ dp, float* __restrict__ dx, int frames, int C, int kernel, int stride,
                   int pooled, long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int c4 = C / 4;
  const int c = static_cast<int>(i % c4) * 4;
  const long long bt = i / c4;
  const int t = static_cast<int>(bt % frames);
  const long long b = bt / frames;
  int j_hi = t / stride;
  if (j_hi > pooled - 1) j_hi = pooled - 1;
  int j_lo = t - kernel + 1 <= 0 ? 0 : (t - kernel + stride) / stride;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = j_lo; j <= j_hi; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(dp + (b * pooled + j) * C + c);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const float inv = 1.0f / kernel;
  *reinterpret_cast<float4*>(dx + bt * C + c) = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
}

// dx[b, u, :] = sum_{j < k, (u - j) % s == 0, t = (u - j) / s < tout} dcol[b, t, j*C + :]
__global__ void __launch_bounds__(256)
col2im_kernel(const __nv_bfloat16* __restrict__ dcol, __nv_bfloat16* __restrict__ dx, int tin, int tout, int k, int s,
              int C, long long total8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = C / 8;
  const int c = static_cast<int>(i % c8) * 8;
  const long long bu = i / c8;
  const int u = static_cast<int>(bu % tin);
  const long long b = bu / tin;
  float acc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) acc[q] = 0.f;
  for (int j = 0; j < k; ++j) {
    const int d = u - j;
    if (d < 0 || d % s != 0) continue;
    const int t = d / s;
    if (t >= tout) continue;
    float v[8];
    ld8bf(dcol + ((b * tout + t) * k + j) * C + c, v);
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] += v[q];
  }
  st8bf(dx + bu * C + c, acc);
}

// ---- conv layer 0 backward: thread = output channel, kR frames per iteration ------------------------------------
constexpr int kC0 = 512, kK0 = 10, kS0 = 5, kR = 8, kW0 = kC0 / 32;

template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float (*s)[N], int warp, int lane) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) s[warp][i] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kW0; ++w) t += s[w][i];
    v[i] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kC0)
conv0_bwd_kernel(const float* __restrict__ wave, long long wave_stride, int samples, int frames, long long total_rows,
                 const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, const __nv_bfloat16* __restrict__ dy, float* __restrict__ dW,
                 float* __restrict__ db, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float s_x[kR][kK0 + 2];
  __shared__ float s_a[kW0][kR];
  __shared__ float s_b[kW0][2 * kR];
  const int c = threadIdx.x, lane = c & 31, warp = c >> 5;
  float wr[kK0], aw[kK0];
#pragma unroll
  for (int j = 0; j < kK0; ++j) {
    wr[j] = w[c * kK0 + j];
    aw[j] = 0.f;
  }
  const float bs = bias[c], gm = gamma[c], bt = beta[c];
  float ab = 0.f, ag = 0.f, abt = 0.f;
  for (long long r0 = static_cast<long long>(blockIdx.x) * kR; r0 < total_rows;
       r0 += static_cast<long long>(gridDim.x) * kR) {
    if (c < kR * kK0) {
      const int rr = c / kK0, j = c - rr * kK0;
      const long long row = r0 + rr;
      float v = 0.f;
      if (row < total_rows) {
        const long long b = row / frames;
        const long long idx = (row - b * frames) * kS0 + j;
        if (idx < samples) v = wave[b * wave_stride + idx];
      }
      s_x[rr][j] = v;
    }
    __syncthreads();
    float pre[kR], d[kR], st[kR];
#pragma unroll
    for (int rr = 0; rr < kR; ++rr) {
      float a = bs;
#pragma unroll
      for (int j = 0; j < kK0; ++j) a = fmaf(wr[j], s_x[rr][j], a);
      pre[rr] = a;
      st[rr] = a;
      d[rr] = (r0 + rr < total_rows) ? __bfloat162float(dy[(r0 + rr) * kC0 + c]) : 0.f;
    }
    block_sum<kR>(st, s_a, warp, lane);
    float mean[kR];
#pragma unroll
    for (int rr = 0; rr < kR; ++rr) {
      mean[rr] = st[rr] * (1.0f / kC0);
      pre[rr] -= mean[rr];
      st[rr] = pre[rr] * pre[rr];
    }
    block_sum<kR>(st, s_a, warp, lane);
    float st2[2 * kR], rstd[kR];
#pragma unroll
    for (int rr = 0; rr < kR; ++rr) {
      rstd[rr] = rsqrtf(st[rr] * (1.0f / kC0) + eps);
      const float xh = pre[rr] * rstd[rr];
      const float dz = d[rr] * gelu_erf_grad(fmaf(gm, xh, bt));
      ag = fmaf(dz, xh, ag);
      abt += dz;
      pre[rr] = xh;
      d[rr] = gm * dz;
      st2[rr] = d[rr];
      st2[kR + rr] = d[rr] * xh;
    }
    block_sum<2 * kR>(st2, s_b, warp, lane);
#pragma unroll
    for (int rr = 0; rr < kR; ++rr) {
      const float dpre = rstd[rr] * (d[rr] - st2[rr] * (1.0f / kC0) - pre[rr] * st2[kR + rr] * (1.0f / kC0));
      ab += dpre;
#pragma unroll
      for (int j = 0; j < kK0; ++j) aw[j] = fmaf(dpre, s_x[rr][j], aw[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < kK0; ++j) atomicAdd(dW + c * kK0 + j, aw[j]);
  atomicAdd(db + c, ab);
  atomicAdd(dgamma + c, ag);
  atomicAdd(dbeta + c, abt);
}

}  // namespace

int layernorm_bwd_ex(const void* x, int x_bf16, const float* gamma, const float* beta, int act_gelu, float eps,
                     const void* dy, int dy_bf16, float* dh, int accumulate, void* dx_bf16, float* dgamma, float* dbeta,
                     long long rows, int C, cudaStream_t stream) {
  B2S_REQUIRE(x && gamma && dy && dgamma && dbeta && (dh || dx_bf16), "layernorm_bwd_ex: null pointer");
  B2S_REQUIRE(!act_gelu || beta != nullptr, "layernorm_bwd_ex: the GELU variant needs beta");
  if (rows <= 0) return B2S_OK;
  long long blocks = (rows + kWarps - 1) / kWarps;
  const long long cap = 4LL * num_sms();
  if (blocks > cap) blocks = cap;
  const unsigned grid = static_cast<unsigned>(blocks);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_bf16);
#define B2S_LNBWD(G)                                                                                                  \
  if (act_gelu)                                                                                                       \
    layernorm_bwd_ex_kernel<G, true><<<grid, kWarps * 32, 0, stream>>>(x, x_bf16, gamma, beta, eps, dy, dy_bf16, dh,  \
                                                                       accumulate, dxb, dgamma, dbeta, rows);        \
  else                                                                                                                \
    layernorm_bwd_ex_kernel<G, false><<<grid, kWarps * 32, 0, stream>>>(x, x_bf16, gamma, beta, eps, dy, dy_bf16, dh, \
                                                                        accumulate, dxb, dgamma, dbeta, rows);
  switch (C) {
    case 256: B2S_LNBWD(1); break;
    case 512: B2S_LNBWD(2); break;
    case 1024: B2S_LNBWD(4); break;
    default:
      set_last_error("layernorm_bwd_ex: unsupported width %d (256, 512, 1024)", C);
      return B2S_ERR_UNSUPPORTED;
  }
#undef B2S_LNBWD
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int colsum_accum(const void* x, int x_bf16, float* out, long long rows, int C, cudaStream_t stream) {
  B2S_REQUIRE(x && out, "colsum_accum: null pointer");
  B2S_REQUIRE(C % 256 == 0, "colsum_accum: C must be a multiple of 256");
  if (rows <= 0) return B2S_OK;
  long long chunks = (rows + kWarps * 16 - 1) / (kWarps * 16);
  const long long cap = (4LL * num_sms() + C / 256 - 1) / (C / 256);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  colsum_kernel<<<dim3(C / 256, static_cast<unsigned>(chunks)), kWarps * 32, 0, stream>>>(x, x_bf16, out, rows, C);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int avgpool_bwd(const float* dpooled, float* dx, int batches, int frames, int C, int kernel, int stride, int pooled,
                cudaStream_t stream) {
  B2S_REQUIRE(dpooled && dx && C % 4 == 0 && kernel > 0 && stride > 0 && pooled > 0, "avgpool_bwd: bad arguments");
  const long long total4 = static_cast<long long>(batches) * frames * (C / 4);
  if (total4 <= 0) return B2S_OK;
  avgpool_bwd_kernel<<<static_cast<unsigned>((total4 + 255) / 256), 256, 0, stream>>>(dpooled, dx, frames, C, kernel,
                                                                                      stride, pooled, total4);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int col2im_add(const void* dcol_bf16, void* dx_bf16, int batches, int tin, int tout, int k, int s, int C,
               cudaStream_t stream) {
  B2S_REQUIRE(dcol_bf16 && dx_bf16 && C % 8 == 0 && k > 0 && s > 0, "col2im_add: bad arguments");
  const long long total8 = static_cast<long long>(batches) * tin * (C / 8);
  if (total8 <= 0) return B2S_OK;
  col2im_kernel<<<static_cast<unsigned>((total8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dcol_bf16), reinterpret_cast<__nv_bfloat16*>(dx_bf16), tin, tout, k, s, C,
      total8);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

int conv0_bwd(const float* wave, long long wave_stride, int batches, int samples, const float* w, const float* bias,
              const float* gamma, const float* beta, float eps, const void* dy_bf16, int frames, float* dW, float* db,
              float* dgamma, float* dbeta, cudaStream_t stream) {
  B2S_REQUIRE(wave && w && bias && gamma && beta && dy_bf16 && dW && db && dgamma && dbeta, "conv0_bwd: null pointer");
  B2S_REQUIRE(frames == (samples - kK0) / kS0 + 1, "conv0_bwd: frames mismatch");
  const long long total = static_cast<long long>(batches) * frames;
  long long blocks = (total + kR - 1) / kR;
  const long long cap = 2LL * num_sms();
  if (blocks > cap) blocks = cap;
  conv0_bwd_kernel<<<static_cast<unsigned>(blocks), kC0, 0, stream>>>(
      wave, wave_stride, samples, frames, total, w, bias, gamma, beta, eps,
      reinterpret_cast<const __nv_bfloat16*>(dy_bf16), dW, db, dgamma, dbeta);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
