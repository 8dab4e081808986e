// Counter-based Exp(1) noise shared by the samplers (ua2_sample.cu, ua2_stream.cu); used only when the caller supplies no
// noise tensor - with a tensor the samplers reproduce torch's own draws.
#pragma once
#include <stdint.h>

namespace ua2 {

// Philox4x32-10 (Salmon et al. 2011) - used only when the caller supplies no noise tensor.
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}
__device__ __forceinline__ float philox_exp1(unsigned long long seed, unsigned long long stream, uint32_t idx) {
  uint32_t c[4] = {idx, (uint32_t)stream, (uint32_t)(stream >> 32), 0x5EEDu};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  const float u = ((float)(c[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
  return -logf(u);
}

}  // namespace ua2
