// gpr_synth.cuh — device generator for synthetic DCGM windows (measurement support).
//
// The reference ships no sample fixtures, so BASELINE.json's configs are
// synthetic by construction.  Recipe (exact constants in DESIGN.md
// §synthetic): every cell is a pure function of (seed, series, t) through the splitmix64
// finaliser, so the GPU, the C oracle and the numpy oracle regenerate identical windows
// without ever moving them.  All values are small integers (or NaN) and exact in f32.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace gpr {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

constexpr uint64_t kTagSeries = 0x5345524945530001ull;
constexpr uint64_t kTagCell = 0x43454C4C00000002ull;
constexpr uint64_t kTagPower = 0x504F574552000003ull;
constexpr uint64_t kTagElig = 0x454C494700000004ull;

// series class: 30 % idle, 10 % single-sample burst, 55 % active, 5 % gappy / young pod
__global__ void k_synth_fill(float* __restrict__ dst, uint64_t seed, int plane, uint64_t series0,
                             uint32_t n_rows, uint32_t T, uint64_t ld) {
  const uint64_t k_series = mix64(seed ^ kTagSeries);
  const uint64_t k_cell = mix64(seed ^ (plane == 0 ? kTagCell : kTagPower));
  for (uint32_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const uint64_t s = series0 + row;
    const uint64_t hs = mix64(k_series ^ s);
    const uint32_t c = (uint32_t)(hs % 100u);
    const uint32_t a = (uint32_t)((hs >> 8) % T);
    const uint64_t b = hs >> 40;
    const bool tail_active = (b & 1ull) != 0;
    float* out = dst + (size_t)row * ld;
    for (uint32_t t = threadIdx.x; t < T; t += blockDim.x) {
      const uint64_t h = mix64(k_cell ^ (s * (uint64_t)T + t));
      float v;
      if (h % 1000u == 0) {
        v = __int_as_float(0x7fc00000);
      } else if (plane == 0) {
        const float v_active = ((h >> 10) & 1ull) ? (float)(1u + (uint32_t)((h >> 11) % 100u)) : 0.0f;
        if (c < 30u) v = 0.0f;
        else if (c < 40u) v = (t == a) ? (float)(1u + (uint32_t)(b % 100u)) : 0.0f;
        else if (c < 95u) v = v_active;
        else v = (t <= a) ? __int_as_float(0x7fc00000) : (tail_active ? v_active : 0.0f);
      } else {
        const bool low = c < 30u || (c >= 95u && !tail_active);
        const uint64_t r = h >> 10;
        v = low ? (float)(40u + (uint32_t)(r % 31u)) : (float)(70u + (uint32_t)(r % 631u));
      }
      out[t] = v;
    }
  }
}

// 5 % of pods fail the Pending / age gates
__global__ void k_synth_eligible(uint8_t* __restrict__ dst, uint64_t seed, uint64_t pod0,
                                 uint32_t n_pods) {
  const uint64_t k = mix64(seed ^ kTagElig);
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n_pods; p += gridDim.x * blockDim.x)
    dst[p] = (uint8_t)((mix64(k ^ (pod0 + p)) % 100u) >= 5u);
}

}  // namespace gpr
