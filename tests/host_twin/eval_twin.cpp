// TEST-ONLY host twin of cosmoslik_b200/csrc/eval_device.cuh (priors + coefficient bytecode + lnl combine):
// the device header itself compiled by g++ -ffp-contract=off with the CUDA qualifiers defined away and the
// round-to-nearest intrinsics mapped to the plain IEEE operations they are.  Built by tests/test_eval_twin.py
// into a temporary directory; never part of the product library.
#include <math.h>
#include <stdint.h>
#include "../../include/cosmoslik_b200.h"
#define CSL_HOST_TWIN
#define __device__
#define __host__
#define __forceinline__ inline
#define CUDART_INF ((double)INFINITY)
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
#include "../../cosmoslik_b200/csrc/eval_device.cuh"

extern "C" {
// x [W, ndim] -> coef [ncoef, ldc], prior [W]; columns W..Wp-1 replay walker W-1 like k_prepare
void twin_prepare(const csl_program_t* p, int W, int Wp, const double* x, double* coef, int ldc, double* prior) {
  for (int w = 0; w < Wp; ++w) {
    const double* xr = x + (long long)(w < W ? w : W - 1) * p->ndim;
    prepare_walker(*p, small_tables_global(*p), xr, w, W, coef, ldc, prior);
  }
}
void twin_combine(const csl_program_t* p, int W, const double* coef, int ldc, const double* prior, const double* chi2,
                  double* lnl) {
  for (int w = 0; w < W; ++w)
    lnl[w] = combine_walker(*p, w, coef, ldc, prior[w], chi2 != nullptr, chi2 != nullptr ? chi2[w] : 0.0);
}
}
