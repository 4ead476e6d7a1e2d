// Per-walker pieces of Slik.evaluate shared by the staged kernels (prepare.cu) and the fused stretch-move
// kernels (samplers.cu): priors + coefficient bytecode on one parameter row, and the final lnl combine.
//   reference: cosmoslik_plugins/likelihoods/priors.py:24-30 ; cosmoslik/cosmoslik.py:135-141
#pragma once
#ifndef CSL_HOST_TWIN      // tests/host_twin/eval_twin.cpp compiles this header with g++ and supplies the few names it needs
#include "common.cuh"
#endif

// The program's small per-walker tables (bytecode, prior boxes, Gaussian priors).  Every thread walks them front
// to back, so read from global memory each table costs a cold L2 / HBM round trip IN SERIES (op -> argument ->
// value, then box index -> bounds, ...): at a few thousand walkers that chain, not arithmetic or bandwidth, was most
// of the 12-17 us of the prepare kernels.  The kernels therefore copy all of them into shared memory with ONE
// round of coalesced loads (stage_small_tables) and interpret from there; programs too long for the staging area
// are interpreted from global memory as before (small_tables_global).
struct csl_small_tables {
  const int32_t* code_op; const int32_t* code_arg; const double* code_val;
  const int32_t* box_idx; const double* box_lo; const double* box_hi;
  const int32_t* gauss_idx; const double* gauss_c; const double* gauss_w;
};
#define CSL_SMALL_STAGE_MAX (16 * 1024)       // bytes of shared memory a kernel spends on the staged tables at most
// bytes of the staged copy: the double tables first, then the int32 ones, padded to 16 bytes
static inline __host__ __device__ int csl_small_tables_bytes(int ncode, int nbox, int ngauss) {
  return ((ncode + 2 * nbox + 2 * ngauss) * 8 + (2 * ncode + nbox + ngauss) * 4 + 15) & ~15;
}
__device__ __forceinline__ csl_small_tables small_tables_global(const csl_program_t& p) {
  csl_small_tables t;
  t.code_op = p.code_op; t.code_arg = p.code_arg; t.code_val = p.code_val;
  t.box_idx = p.box_idx; t.box_lo = p.box_lo; t.box_hi = p.box_hi;
  t.gauss_idx = p.gauss_idx; t.gauss_c = p.gauss_c; t.gauss_w = p.gauss_w;
  return t;
}
#ifndef CSL_HOST_TWIN
// all threads of the CTA call this; the copy is complete after the caller's next __syncthreads()
__device__ __forceinline__ csl_small_tables stage_small_tables(const csl_program_t& p, double* sm, int tid, int nthr) {
  double* dv = sm;                       double* dlo = dv + p.ncode;   double* dhi = dlo + p.nbox;
  double* dc = dhi + p.nbox;             double* dw = dc + p.ngauss;
  int32_t* iop = reinterpret_cast<int32_t*>(dw + p.ngauss);
  int32_t* iarg = iop + p.ncode;         int32_t* ibox = iarg + p.ncode;   int32_t* igau = ibox + p.nbox;
  for (int i = tid; i < p.ncode; i += nthr) { dv[i] = p.code_val[i]; iop[i] = p.code_op[i]; iarg[i] = p.code_arg[i]; }
  for (int i = tid; i < p.nbox; i += nthr) { dlo[i] = p.box_lo[i]; dhi[i] = p.box_hi[i]; ibox[i] = p.box_idx[i]; }
  for (int i = tid; i < p.ngauss; i += nthr) { dc[i] = p.gauss_c[i]; dw[i] = p.gauss_w[i]; igau[i] = p.gauss_idx[i]; }
  csl_small_tables t;
  t.code_op = iop; t.code_arg = iarg; t.code_val = dv;
  t.box_idx = ibox; t.box_lo = dlo; t.box_hi = dhi;
  t.gauss_idx = igau; t.gauss_c = dc; t.gauss_w = dw;
  return t;
}
#endif

// xr: the walker's parameter row (any address space); w: its column; columns w >= W are tile padding
// (their coefficients are written, their prior is not).  t: the small tables (shared or global memory).
__device__ __forceinline__ void prepare_walker(const csl_program_t& p, const csl_small_tables& t, const double* xr, int w,
                                               int W, double* __restrict__ coef, int ldc, double* __restrict__ prior) {
  // coefficient bytecode first: priors may refer to derived quantities (coefficient rows)
  double st[CSL_STACK];
  int sp = 0;
  for (int pc = 0; pc < p.ncode; ++pc) {
    const int op = t.code_op[pc];
    switch (op) {
      case CSL_OP_PUSHP: st[sp++] = xr[t.code_arg[pc]]; break;
      case CSL_OP_PUSHC: st[sp++] = t.code_val[pc]; break;
      case CSL_OP_ADD: --sp; st[sp - 1] = __dadd_rn(st[sp - 1], st[sp]); break;
      case CSL_OP_SUB: --sp; st[sp - 1] = __dsub_rn(st[sp - 1], st[sp]); break;
      case CSL_OP_MUL: --sp; st[sp - 1] = __dmul_rn(st[sp - 1], st[sp]); break;
      case CSL_OP_DIV: --sp; st[sp - 1] = __ddiv_rn(st[sp - 1], st[sp]); break;
      case CSL_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
      case CSL_OP_EXP: st[sp - 1] = exp(st[sp - 1]); break;
      case CSL_OP_LOG: st[sp - 1] = log(st[sp - 1]); break;
      case CSL_OP_POW: --sp; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
      case CSL_OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
      case CSL_OP_SQR: st[sp - 1] = __dmul_rn(st[sp - 1], st[sp - 1]); break;
      case CSL_OP_STORE: coef[(long long)t.code_arg[pc] * ldc + w] = st[--sp]; break;
      default: break;
    }
  }
  // A prior's subject is a sampled parameter (index >= 0) or a DERIVED quantity the script sets in __call__
  // (index < 0: coefficient row -index - 1; the reference evaluates priors again after the call for exactly
  // these, cosmoslik.py:139-141).  This thread wrote that coefficient a few lines above.
  // priors.py:26-28: first box that fails `lo <= x <= hi` (NaN fails) gives +inf
  bool out = false;
  for (int b = 0; b < p.nbox; ++b) {
    const int ix = t.box_idx[b];
    double v = ix >= 0 ? xr[ix] : coef[(long long)(-ix - 1) * ldc + w];
    if (!(t.box_lo[b] <= v && v <= t.box_hi[b])) out = true;
  }
  // priors.py:30: sum((x-c)**2./2/w**2), python sum from 0 in list order
  double pr = 0.0;
  for (int k = 0; k < p.ngauss; ++k) {
    const int ix = t.gauss_idx[k];
    double d = __dsub_rn(ix >= 0 ? xr[ix] : coef[(long long)(-ix - 1) * ldc + w], t.gauss_c[k]);
    double wv = t.gauss_w[k];
    pr = __dadd_rn(pr, __ddiv_rn(__ddiv_rn(__dmul_rn(d, d), 2.0), __dmul_rn(wv, wv)));
  }
  if (w < W) prior[w] = out ? CUDART_INF : pr;
}

// lnl = like + prior with like = sum scalar coefs (+ chi2/2), cosmoslik.py:141; +inf on a hard prior (:135-137)
__device__ __forceinline__ double combine_walker(const csl_program_t& p, int w, const double* __restrict__ coef, int ldc,
                                                 double pr, bool has_chi2, double chi2) {
  double like = 0.0;
  bool first = true;
  for (int k = 0; k < p.nscalar; ++k) {
    double v = coef[(long long)p.scalar_coef[k] * ldc + w];
    like = first ? v : __dadd_rn(like, v);
    first = false;
  }
  if (has_chi2) {
    double h = chi2 * 0.5;
    like = first ? h : __dadd_rn(like, h);
  }
  return (pr == CUDART_INF) ? CUDART_INF : __dadd_rn(like, pr);
}
