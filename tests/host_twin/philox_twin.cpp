// TEST-ONLY host twin of cosmoslik_b200/csrc/philox.cuh: the device header itself is compiled by g++ with
// the CUDA qualifiers and the two intrinsics it uses defined away, so that the CPU suite can check the
// very source the kernels include against the Random123 known answers and oracle/philox.py.
// Built by tests/test_philox_twin.py into a temporary directory; never part of the product library.
#include <math.h>
#include <stdint.h>
#define __device__
#define __forceinline__ inline
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
#include "../../cosmoslik_b200/csrc/philox.cuh"

extern "C" {
void twin_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  philox_out o = philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = o.w[i];
}
double twin_u01(uint32_t hi, uint32_t lo) { return philox_u01(hi, lo); }
void twin_mh(uint64_t seed, uint32_t step, uint32_t chain0, int W, int nd, double* z, double* u) {
  for (int w = 0; w < W; ++w) {
    for (int p = 0; 2 * p < nd; ++p) {
      double a, b;
      philox_normal_pair(seed, chain0 + (uint32_t)w, (uint32_t)p, step, a, b);
      z[(long long)w * nd + 2 * p] = a;
      if (2 * p + 1 < nd) z[(long long)w * nd + 2 * p + 1] = b;
    }
    u[w] = philox_mh_uniform(seed, chain0 + (uint32_t)w, step);
  }
}
void twin_stretch(uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0, int Ns, uint32_t Nc, double* uz,
                  int64_t* j, double* ua) {
  for (int k = 0; k < Ns; ++k) philox_stretch(seed, walker0 + (uint32_t)k, half, step, Nc, uz[k], j[k], ua[k]);
}
}
