// Counter-based random numbers for the kernels that draw their own noise (lm_noise.cu, lm_rollout.cu): Philox4x32-10 keyed by a
// seed, addressed by (stream position, element index) - no generator state in memory, no dependence on the launch geometry.
#pragma once
#include "lm_common.cuh"

namespace lm {

// ---- Philox4x32-10 (Salmon et al., SC'11): counter = (element, stream), key = seed
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned int hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned int hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
// 128 random bits of element `element` at stream position `stream` under `seed`
__device__ __forceinline__ uint4 philox_at(unsigned long long seed, unsigned long long stream, unsigned long long element) {
  return philox4x32_10(make_uint4((unsigned int)element, (unsigned int)(element >> 32), (unsigned int)stream, (unsigned int)(stream >> 32)),
                       make_uint2((unsigned int)seed, (unsigned int)(seed >> 32)));
}
// four U[0,1) draws (24-bit mantissas, as torch.rand)
__device__ __forceinline__ float4 uniform4(const uint4& r) {
  return make_float4((float)(r.x >> 8) * 5.9604644775390625e-08f, (float)(r.y >> 8) * 5.9604644775390625e-08f,
                     (float)(r.z >> 8) * 5.9604644775390625e-08f, (float)(r.w >> 8) * 5.9604644775390625e-08f);
}
// four N(0,1) draws: Box-Muller on two pairs
__device__ __forceinline__ float4 normal4(const uint4& r) {
  const float u0 = ((float)(r.x >> 8) + 1.0f) * 5.9604644775390625e-08f, u1 = ((float)(r.z >> 8) + 1.0f) * 5.9604644775390625e-08f;   // (0, 1]
  const float r0 = sqrtf(-2.0f * logf(u0)), r1 = sqrtf(-2.0f * logf(u1));
  float s0, c0, s1, c1;
  sincospif((float)r.y * 4.656612873077393e-10f, &s0, &c0);       // 2 pi * (y / 2^32)
  sincospif((float)r.w * 4.656612873077393e-10f, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

}  // namespace lm
