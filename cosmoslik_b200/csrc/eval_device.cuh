// Per-walker pieces of Slik.evaluate shared by the staged kernels (prepare.cu) and the fused stretch-move
// kernels (samplers.cu): priors + coefficient bytecode on one parameter row, and the final lnl combine.
//   reference: cosmoslik_plugins/likelihoods/priors.py:24-30 ; cosmoslik/cosmoslik.py:135-141
#pragma once
#ifndef CSL_HOST_TWIN      // tests/host_twin/eval_twin.cpp compiles this header with g++ and supplies the few names it needs
#include "common.cuh"
#endif

// xr: the walker's parameter row (any address space); w: its column; columns w >= W are tile padding
// (their coefficients are written, their prior is not).
__device__ __forceinline__ void prepare_walker(const csl_program_t& p, const double* xr, int w, int W,
                                               double* __restrict__ coef, int ldc, double* __restrict__ prior) {
  // coefficient bytecode first: priors may refer to derived quantities (coefficient rows)
  double st[CSL_STACK];
  int sp = 0;
  for (int pc = 0; pc < p.ncode; ++pc) {
    const int op = p.code_op[pc];
    switch (op) {
      case CSL_OP_PUSHP: st[sp++] = xr[p.code_arg[pc]]; break;
      case CSL_OP_PUSHC: st[sp++] = p.code_val[pc]; break;
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
      case CSL_OP_STORE: coef[(long long)p.code_arg[pc] * ldc + w] = st[--sp]; break;
      default: break;
    }
  }
  // A prior's subject is a sampled parameter (index >= 0) or a DERIVED quantity the script sets in __call__
  // (index < 0: coefficient row -index - 1; the reference evaluates priors again after the call for exactly
  // these, cosmoslik.py:139-141).  This thread wrote that coefficient a few lines above.
  // priors.py:26-28: first box that fails `lo <= x <= hi` (NaN fails) gives +inf
  bool out = false;
  for (int b = 0; b < p.nbox; ++b) {
    const int ix = p.box_idx[b];
    double v = ix >= 0 ? xr[ix] : coef[(long long)(-ix - 1) * ldc + w];
    if (!(p.box_lo[b] <= v && v <= p.box_hi[b])) out = true;
  }
  // priors.py:30: sum((x-c)**2./2/w**2), python sum from 0 in list order
  double pr = 0.0;
  for (int k = 0; k < p.ngauss; ++k) {
    const int ix = p.gauss_idx[k];
    double d = __dsub_rn(ix >= 0 ? xr[ix] : coef[(long long)(-ix - 1) * ldc + w], p.gauss_c[k]);
    double wv = p.gauss_w[k];
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
