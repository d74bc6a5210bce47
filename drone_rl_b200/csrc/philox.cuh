// Counter-based Philox4x32-10 (Salmon et al., SC'11).  Stream convention shared with the
// test oracle (oracle/philox.py): key = 64-bit seed, counter = (env_id lo, env_id hi, index,
// stream | index_hi << 8).  The reference draws from numpy's global MT19937
// (/root/reference/drone.py:57,73); parity is defined by feeding these uniforms into it.
#pragma once
#include <stdint.h>

namespace dronecu {

enum : uint32_t { STREAM_RESET = 0, STREAM_ACTION = 2, STREAM_NOISE = 3 };

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
    k0 += W0;
    k1 += W1;
  }
  return c;
}

// Round keys of a 64-bit seed (k0 + r*W0, k1 + r*W1), computed once on the host and kept in the kernel
// parameter (constant) bank: the ten key bumps per call leave the instruction stream (ncu, round 1: 18
// UIADD3 per call in an issue-bound kernel).
struct PhiloxKeys {
  uint32_t k[10][2];
};

inline void philox_expand_key(uint64_t seed, PhiloxKeys& K) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; ++r) {
    K.k[r][0] = k0; K.k[r][1] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys& K) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ K.k[r][0], lo1, hi0 ^ c.w ^ K.k[r][1], lo0);
  }
  return c;
}

// 24 random bits -> [0,1): exactly representable in float32 and float64 alike.
__device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * 5.9604644775390625e-8f; }

__device__ __forceinline__ uint4 env_stream(uint64_t seed, uint64_t env_id, uint64_t index, uint32_t stream) {
  const uint4 c = make_uint4((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)index,
                             stream | ((uint32_t)(index >> 32) << 8));
  return philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
}

__device__ __forceinline__ uint4 env_stream(const PhiloxKeys& K, uint64_t env_id, uint64_t index, uint32_t stream) {
  const uint4 c = make_uint4((uint32_t)env_id, (uint32_t)(env_id >> 32), (uint32_t)index,
                             stream | ((uint32_t)(index >> 32) << 8));
  return philox4x32_10(c, K);
}

}  // namespace dronecu
