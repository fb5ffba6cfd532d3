// rng.cuh -- counter-based keep/drop decisions for the train-mode regularisers of the HuBERT encoder
// (nn.Dropout sites of TF/models/hubert/modeling_hubert.py:223-230,351-368,383-393,557-587 and the attention-probability
// dropout :254). Nothing is stored: forward and backward regenerate the same decision for an element from
// (seed, site, element index), so the backward needs no mask tensors in HBM.
//
//   mix(x, k)      : x ^= x >> 16; x = x * 0x7feb352d + k; x ^= x >> 15; x *= 0x846ca68b; x ^= x >> 16     (uint32)
//   stream key     : t = mix(site ^ seed_lo, seed_hi); t = mix(t ^ a, 0x9E3779B9); t = mix(t ^ b, 0x85EBCA6B);
//                    k1 = t; k2 = mix(t ^ seed_hi, seed_lo)                (a, b = 0 for elementwise sites;
//                                                                           a = sequence, b = head for attention)
//   element kept  <=>  mix(elem ^ k1, k2) >= thresh,   thresh = floor(p * 2^32); kept elements are scaled by 1/(1-p)
//   elem           : row * N + col of the [rows, N] activation (elementwise sites), (query << 16) | key (attention)
//
// oracle/regularizers.py restates exactly this in numpy so the CPU oracle can be run with the same masks.
#pragma once
#include <cstdint>

namespace b2s {

enum DropSite : uint32_t {
  SITE_FEAT_PROJ = 1,   // HubertFeatureProjection.dropout
  SITE_POS_ADD = 2,     // HubertEncoderStableLayerNorm.dropout (after hidden + positional conv)
  SITE_LAYER0 = 16,     // per layer l: SITE_LAYER0 + 4*l + {0: attention output, 1: FFN activation, 2: FFN output,
                        //                                   3: attention probabilities}
};
__host__ __device__ inline uint32_t site_attn_out(int l) { return SITE_LAYER0 + 4u * l; }
__host__ __device__ inline uint32_t site_ff_act(int l) { return SITE_LAYER0 + 4u * l + 1u; }
__host__ __device__ inline uint32_t site_ff_out(int l) { return SITE_LAYER0 + 4u * l + 2u; }
__host__ __device__ inline uint32_t site_attn_prob(int l) { return SITE_LAYER0 + 4u * l + 3u; }

__host__ __device__ __forceinline__ uint32_t rng_mix(uint32_t x, uint32_t k) {
  x ^= x >> 16;
  x = x * 0x7feb352du + k;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}

__host__ __device__ inline void rng_stream_key(unsigned long long seed, uint32_t site, uint32_t a, uint32_t b,
                                               uint32_t* k1, uint32_t* k2) {
  const uint32_t s0 = static_cast<uint32_t>(seed), s1 = static_cast<uint32_t>(seed >> 32);
  uint32_t t = rng_mix(site ^ s0, s1);
  t = rng_mix(t ^ a, 0x9E3779B9u);
  t = rng_mix(t ^ b, 0x85EBCA6Bu);
  *k1 = t;
  *k2 = rng_mix(t ^ s1, s0);
}

// one elementwise dropout site, resolved on the host
struct DropSpec {
  uint32_t k1, k2;
  uint32_t thresh;  // 0 = dropout off
  float inv_keep;
};

// attention-probability dropout of one launch; the stream key depends on (sequence, head) and is derived per CTA
struct AttnDrop {
  unsigned long long seed;
  uint32_t site;
  uint32_t thresh;
  float inv_keep;
};

inline uint32_t drop_threshold(float p) {
  if (!(p > 0.f)) return 0u;
  const double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 4294967295u : static_cast<uint32_t>(t);
}

inline DropSpec make_drop_spec(unsigned long long seed, uint32_t site, float p) {
  DropSpec d{};
  d.thresh = drop_threshold(p);
  d.inv_keep = p < 1.f ? 1.0f / (1.0f - p) : 0.f;
  rng_stream_key(seed, site, 0u, 0u, &d.k1, &d.k2);
  return d;
}

__device__ __forceinline__ bool rng_keep(uint32_t elem, uint32_t k1, uint32_t k2, uint32_t thresh) {
  return rng_mix(elem ^ k1, k2) >= thresh;
}

}  // namespace b2s
