// logmel.cu -- Whisper log-mel feature extraction on the GPU (the step before the path).
// The reference runs transformers' WhisperFeatureExtractor on the CPU inside its collate function
// (REF/trainer.py:178-182; TF/models/whisper/feature_extraction_whisper.py:105-133): reflect-padded STFT with a
// periodic Hann window (n_fft 400, hop 160), power spectrum, 80 slaney mel filters, log10(max(., 1e-10)), clamp to
// (max over the utterance - 8), (x + 4) / 4, last frame dropped. At hundreds of utterances per second per GPU that
// numpy STFT is the bottleneck, so the same arithmetic runs here in fp32:
//   logmel_power_kernel : kFrames frames per block; thread k = DFT bin k (0..200) over the windowed frame held in
//                         shared memory (direct 400-point DFT from a twiddle table: 0.5 GFLOP per 30 s utterance),
//                         then thread m = mel bin m; writes log10 mel energies [B, 80, frames] and the per-utterance
//                         maximum (ordered-int atomicMax).
//   logmel_finalize_kernel : x = (max(x, max_b - 8) + 4) / 4.
#include "b2s_common.cuh"
#include "ops.cuh"

namespace b2s {
namespace {

constexpr int kNfft = 400, kHop = 160, kBins = 201, kMels = 80, kFrames = 8;

__device__ __forceinline__ int float_to_ordered(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void __launch_bounds__(256)
logmel_power_kernel(const float* __restrict__ wave, long long wave_stride, int samples, int frames,
                    const float* __restrict__ mel /* [201, 80] */, float* __restrict__ out /* [B, 80, frames] */,
                    int* __restrict__ max_ord /* [B] */) {
  __shared__ float s_tw[2][kNfft];          // cos / sin of 2 pi j / 400
  __shared__ float s_x[kFrames][kNfft];     // windowed frames
  __shared__ float s_p[kFrames][kBins + 3]; // power spectra
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFrames;
  const float* x = wave + static_cast<long long>(b) * wave_stride;
  for (int j = threadIdx.x; j < kNfft; j += blockDim.x) {
    float sn, cs;
    sincospif(2.0f * j / kNfft, &sn, &cs);
    s_tw[0][j] = cs;
    s_tw[1][j] = sn;
  }
  for (int i = threadIdx.x; i < kFrames * kNfft; i += blockDim.x) {
    const int f = i / kNfft, n = i - f * kNfft;
    int idx = (t0 + f) * kHop + n - kNfft / 2;  // center = True, reflect padding
    if (idx < 0) idx = -idx;
    if (idx >= samples) idx = 2 * (samples - 1) - idx;
    const float win = 0.5f - 0.5f * cospif(2.0f * n / kNfft);  // periodic Hann
    s_x[f][n] = (t0 + f < frames && idx >= 0 && idx < samples) ? x[idx] * win : 0.f;
  }
  __syncthreads();
  const int k = threadIdx.x;
  if (k < kBins) {
    float re[kFrames], im[kFrames];
#pragma unroll
    for (int f = 0; f < kFrames; ++f) re[f] = im[f] = 0.f;
    int tw = 0;
    for (int n = 0; n < kNfft; ++n) {
      const float cs = s_tw[0][tw], sn = s_tw[1][tw];
#pragma unroll
      for (int f = 0; f < kFrames; ++f) {
        const float v = s_x[f][n];
        re[f] = fmaf(v, cs, re[f]);
        im[f] = fmaf(v, sn, im[f]);
      }
      tw += k;
      if (tw >= kNfft) tw -= kNfft;
    }
#pragma unroll
    for (int f = 0; f < kFrames; ++f) s_p[f][k] = re[f] * re[f] + im[f] * im[f];
  }
  __syncthreads();
  float local_max = -INFINITY;
  for (int i = threadIdx.x; i < kFrames * kMels; i += blockDim.x) {
    const int m = i / kFrames, f = i - m * kFrames;  // consecutive threads -> consecutive frames of one mel bin
    if (t0 + f >= frames) continue;
    float acc = 0.f;
    for (int kk = 0; kk < kBins; ++kk) acc = fmaf(__ldg(mel + kk * kMels + m), s_p[f][kk], acc);
    const float v = log10f(fmaxf(acc, 1e-10f));
    out[(static_cast<long long>(b) * kMels + m) * frames + t0 + f] = v;
    local_max = fmaxf(local_max, v);
  }
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0 && local_max > -INFINITY) atomicMax(max_ord + b, float_to_ordered(local_max));
}

__global__ void __launch_bounds__(256)
logmel_finalize_kernel(float* __restrict__ out, const int* __restrict__ max_ord, long long per_utt, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float mx = ordered_to_float(max_ord[i / per_utt]);
  out[i] = (fmaxf(out[i], mx - 8.0f) + 4.0f) * 0.25f;
}

__global__ void fill_int_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace

int whisper_log_mel(const float* wave, long long wave_stride, int batches, int samples, const float* mel_filters,
                    float* out, int frames, int* max_ws, cudaStream_t stream) {
  B2S_REQUIRE(wave && mel_filters && out && max_ws, "whisper_log_mel: null pointer");
  B2S_REQUIRE(batches > 0 && samples > kNfft / 2 && frames == samples / kHop,
              "whisper_log_mel: frames must equal samples / %d (the extractor drops the last STFT frame)", kHop);
  fill_int_kernel<<<(batches + 255) / 256, 256, 0, stream>>>(max_ws, batches, static_cast<int>(0x80000000u));
  B2S_LAUNCH_CHECK();
  logmel_power_kernel<<<dim3((frames + kFrames - 1) / kFrames, batches), 256, 0, stream>>>(wave, wave_stride, samples,
                                                                                         frames, mel_filters, out, max_ws);
  B2S_LAUNCH_CHECK();
  const long long per_utt = static_cast<long long>(kMels) * frames, total = per_utt * batches;
  logmel_finalize_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(out, max_ws, per_utt, total);
  B2S_LAUNCH_CHECK();
  return B2S_OK;
}

}  // namespace b2s
