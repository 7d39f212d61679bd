// Counter-based Philox4x32-10 generator shared by the sampler (Gaussian noise drawn in registers) and the dropout masks of
// the training-mode GroupNorm kernels (recomputed, never stored, in the backward pass).
#pragma once
#include <stdint.h>

namespace {

struct Philox {
  static __device__ __forceinline__ uint4 round10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
      const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
      ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
      key.x += W0;
      key.y += W1;
    }
    return ctr;
  }
};


// dropout keep-scale for the 4 channels of element quad `q` of stream `stream`: 0 or 1/(1-p)   (nn.Dropout, models/layerspp.py:278)
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint32_t stream, uint64_t q, float p) {
  const uint4 r = Philox::round10(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream, 0x44524F50u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t thr = (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);
  const float k = 1.0f / (1.0f - p);
  return make_float4(r.x >= thr ? k : 0.f, r.y >= thr ? k : 0.f, r.z >= thr ? k : 0.f, r.w >= thr ? k : 0.f);
}

}  // namespace
