// Priors + coefficient bytecode for W walkers, and the final lnl combine.
//   reference: cosmoslik_plugins/likelihoods/priors.py:24-30 ; cosmoslik/cosmoslik.py:135-141
#include "common.cuh"

#define PREP_THREADS 128

// One thread per walker.  The x tile [128, ndim] is staged through shared memory with
// coalesced loads (row stride ndim+1 to spread banks), then each thread interprets the
// bytecode on its own row.  HBM traffic: 8*ndim B read + 8*(ncoef+1) B written per walker.
__global__ void __launch_bounds__(PREP_THREADS)
k_prepare(csl_program_t p, int W, int Wp, const double* __restrict__ x, double* __restrict__ coef, int ldc,
          double* __restrict__ prior) {
  extern __shared__ double sx[];
  const int nd = p.ndim, lds = nd + 1;
  const int w0 = blockIdx.x * PREP_THREADS;
  const int nw = min(PREP_THREADS, Wp - w0);
  // columns W..Wp-1 (padding up to the 128-walker tile) replay walker W-1 so they stay finite
  for (int i = threadIdx.x; i < nw * nd; i += PREP_THREADS) {
    int r = i / nd, c = i - r * nd;
    sx[r * lds + c] = x[(long long)min(w0 + r, W - 1) * nd + c];
  }
  __syncthreads();
  if (threadIdx.x >= nw) return;
  const double* xr = sx + threadIdx.x * lds;
  const int w = w0 + threadIdx.x;

  // priors.py:26-28: first box that fails `lo <= x <= hi` (NaN fails) gives +inf
  bool out = false;
  for (int b = 0; b < p.nbox; ++b) {
    double v = xr[p.box_idx[b]];
    if (!(p.box_lo[b] <= v && v <= p.box_hi[b])) out = true;
  }
  // priors.py:30: sum((x-c)**2./2/w**2), python sum from 0 in list order
  double pr = 0.0;
  for (int k = 0; k < p.ngauss; ++k) {
    double d = __dsub_rn(xr[p.gauss_idx[k]], p.gauss_c[k]);
    double wv = p.gauss_w[k];
    pr = __dadd_rn(pr, __ddiv_rn(__ddiv_rn(__dmul_rn(d, d), 2.0), __dmul_rn(wv, wv)));
  }
  if (w < W) prior[w] = out ? CUDART_INF : pr;

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
}

// lnl = like + prior with like = sum scalar coefs (+ chi2/2), cosmoslik.py:141; +inf on a hard prior (:135-137)
__global__ void k_combine(csl_program_t p, int W, const double* __restrict__ coef, int ldc,
                          const double* __restrict__ prior, const double* __restrict__ chi2,
                          double* __restrict__ lnl) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double pr = prior[w];
  double like = 0.0;
  bool first = true;
  for (int k = 0; k < p.nscalar; ++k) {
    double v = coef[(long long)p.scalar_coef[k] * ldc + w];
    like = first ? v : __dadd_rn(like, v);
    first = false;
  }
  if (chi2 != nullptr) {
    double h = chi2[w] * 0.5;
    like = first ? h : __dadd_rn(like, h);
  }
  lnl[w] = (pr == CUDART_INF) ? CUDART_INF : __dadd_rn(like, pr);
}

// resid_mode 2: R[i, w] = coef[direct_coef[i], w] - data[i]; rows n..n_pad are zero
__global__ void k_direct_residual(csl_program_t p, int W, const double* __restrict__ coef, int ldc,
                                  double* __restrict__ R, int ldr) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (w >= ldr) return;
  double v = 0.0;
  if (i < p.n && w < W) v = __dsub_rn(coef[(long long)p.direct_coef[i] * ldc + w], p.data[i]);
  R[(long long)i * ldr + w] = v;
}

extern "C" int csl_eval_prepare(const csl_program_t* p, int W, const double* x, double* coef, int ldc,
                                double* prior, void* stream) {
  CSL_REQUIRE(p && x && prior && W > 0, "csl_eval_prepare: null argument or W <= 0");
  CSL_REQUIRE(p->ndim >= 1 && p->ndim <= CSL_MAX_DIM, "csl_eval_prepare: ndim out of range");
  CSL_REQUIRE(p->ncoef == 0 || (coef && ldc >= W), "csl_eval_prepare: coef buffer / ldc");
  const int Wp = (p->ncoef > 0) ? ldc : W;   // coefficient columns W..ldc-1 replay walker W-1
  int grid = (Wp + PREP_THREADS - 1) / PREP_THREADS;
  size_t smem = (size_t)PREP_THREADS * (p->ndim + 1) * sizeof(double);
  k_prepare<<<grid, PREP_THREADS, smem, (cudaStream_t)stream>>>(*p, W, Wp, x, coef, ldc, prior);
  return csl_check_launch("k_prepare");
}

extern "C" int csl_combine(const csl_program_t* p, int W, const double* coef, int ldc, const double* prior,
                           const double* chi2, double* lnl, void* stream) {
  CSL_REQUIRE(p && prior && lnl && W > 0, "csl_combine: null argument");
  k_combine<<<(W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, W, coef, ldc, prior, chi2, lnl);
  return csl_check_launch("k_combine");
}

extern "C" int csl_direct_residual(const csl_program_t* p, int W, const double* coef, int ldc, double* R,
                                   int ldr, void* stream) {
  CSL_REQUIRE(p && coef && R && W > 0 && ldr >= W, "csl_direct_residual: bad argument");
  CSL_REQUIRE(p->resid_mode == 2, "csl_direct_residual: program has no direct residual");
  dim3 grid((ldr + 255) / 256, p->n_pad);
  k_direct_residual<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, W, coef, ldc, R, ldr);
  return csl_check_launch("k_direct_residual");
}
