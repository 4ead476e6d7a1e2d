// Philox4x32-10 counter-based streams (Salmon et al. SC'11); layout documented in
// oracle/philox.py, which is the numpy statement these device functions are tested against.
#pragma once
#include <stdint.h>

#define CSL_STREAM_MH_Z 1u
#define CSL_STREAM_MH_U 2u
#define CSL_STREAM_STRETCH 3u
#define CSL_STREAM_PARTNER 4u

struct philox_out { uint32_t w[4]; };

__device__ __forceinline__ philox_out philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  philox_out o; o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
  return o;
}

// two words -> double in (0, 1]: never 0 (log is taken of it); for m >= 2^52 the +0.5 is absorbed by rounding,
// and the single value m = 2^53 - 1 (probability 2^-53) gives exactly 1.0
__device__ __forceinline__ double philox_u01(uint32_t hi, uint32_t lo) {
  uint64_t m = ((uint64_t)(hi >> 5) << 26) + (uint64_t)(lo >> 6);
  return ((double)m + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ void philox_normal_pair(uint64_t seed, uint32_t chain, uint32_t pair, uint32_t step,
                                                   double& z0, double& z1) {
  philox_out o = philox4x32_10(chain, pair, step, CSL_STREAM_MH_Z, (uint32_t)seed, (uint32_t)(seed >> 32));
  double u1 = philox_u01(o.w[0], o.w[1]), u2 = philox_u01(o.w[2], o.w[3]);
  double r = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = r * c; z1 = r * s;
}

__device__ __forceinline__ double philox_mh_uniform(uint64_t seed, uint32_t chain, uint32_t step) {
  philox_out o = philox4x32_10(chain, 0u, step, CSL_STREAM_MH_U, (uint32_t)seed, (uint32_t)(seed >> 32));
  return philox_u01(o.w[0], o.w[1]);
}

__device__ __forceinline__ void philox_stretch(uint64_t seed, uint32_t walker, uint32_t half, uint32_t step,
                                               uint32_t ncomp, double& u_z, int64_t& j, double& u_acc) {
  philox_out o = philox4x32_10(walker, half, step, CSL_STREAM_STRETCH, (uint32_t)seed, (uint32_t)(seed >> 32));
  philox_out v = philox4x32_10(walker, half, step, CSL_STREAM_PARTNER, (uint32_t)seed, (uint32_t)(seed >> 32));
  u_z = philox_u01(o.w[0], o.w[1]);
  u_acc = philox_u01(o.w[2], o.w[3]);
  j = (int64_t)(((uint64_t)v.w[0] * (uint64_t)ncomp) >> 32);
}
