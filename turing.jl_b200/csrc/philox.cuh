// philox.cuh -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), device side.
// key = 64-bit seed, counter = (global chain id, iteration, slot, sub).  Every draw of the
// sampler is addressed, never streamed, so chains are independent of thread/block/GPU
// placement: "same seed => same chains" (test/mcmc/hmc.jl:164-171) holds for any sharding.
#pragma once
#include "common.cuh"

namespace tnuts {

// NOT inlined on purpose: the thread-per-chain kernel draws from ~25 call sites; inlining the
// 10 rounds everywhere made its SASS 133 KB (instruction-fetch stalls in the ncu profile)
static __device__ __noinline__ uint4 philox4x32_10(unsigned long long seed, uint32_t c0, uint32_t c1,
                                                   uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0;
    const uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0;
    c1 = lo1;
    c2 = n2;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
  const unsigned long long b = (((unsigned long long)hi << 32) | lo) >> 11;
  return ((double)b + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void uniform2(unsigned long long seed, uint32_t chain, uint32_t iter,
                                         uint32_t slot, uint32_t sub, double &u0, double &u1) {
  const uint4 o = philox4x32_10(seed, chain, iter, slot, sub);
  u0 = u53(o.x, o.y);
  u1 = u53(o.z, o.w);
}

__device__ __forceinline__ double uniform1(unsigned long long seed, uint32_t chain, uint32_t iter,
                                           uint32_t slot, uint32_t sub) {
  const uint4 o = philox4x32_10(seed, chain, iter, slot, sub);
  return u53(o.x, o.y);
}

// Box-Muller: two normals per Philox block
static __device__ __noinline__ void normal2(unsigned long long seed, uint32_t chain, uint32_t iter,
                                        uint32_t slot, uint32_t sub, double &z0, double &z1) {
  double u0, u1;
  uniform2(seed, chain, iter, slot, sub, u0, u1);
  const double rad = sqrt(-2.0 * log(u0));
  double s, c;
  sincospi(2.0 * u1, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

}  // namespace tnuts
