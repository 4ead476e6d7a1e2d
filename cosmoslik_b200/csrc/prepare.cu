// Priors + coefficient bytecode for W walkers, and the final lnl combine.
//   reference: cosmoslik_plugins/likelihoods/priors.py:24-30 ; cosmoslik/cosmoslik.py:135-141
#include "eval_device.cuh"

#define PREP_THREADS 128

// One thread per walker.  The x tile [128, ndim] is staged through shared memory with
// coalesced loads (row stride ndim+1 to spread banks), then each thread interprets the
// bytecode on its own row.  HBM traffic: 8*ndim B read + 8*(ncoef+1) B written per walker.
__global__ void __launch_bounds__(PREP_THREADS)
k_prepare(csl_program_t p, int W, int Wp, const double* __restrict__ x, double* __restrict__ coef, int ldc,
          double* __restrict__ prior, int stage_tables) {
  extern __shared__ __align__(16) double sx[];
  const int nd = p.ndim, lds = nd + 1;
  const int w0 = blockIdx.x * PREP_THREADS;
  const int nw = min(PREP_THREADS, Wp - w0);
  // the bytecode and prior tables ride in with the x tile: one round of loads, one barrier
  const csl_small_tables tabs = stage_tables ? stage_small_tables(p, sx + PREP_THREADS * lds, threadIdx.x, PREP_THREADS)
                                             : small_tables_global(p);
  // columns W..Wp-1 (padding up to the 128-walker tile) replay walker W-1 so they stay finite
  // four elements per thread and trip: the loads of a trip are independent and leave together (one element per
  // trip made every thread wait out nd / 128 * 128 global-memory latencies in series)
  const int tot = nw * nd;
  for (int i0 = threadIdx.x; i0 < tot; i0 += 4 * PREP_THREADS) {
    double v[4];
    int dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * PREP_THREADS;
      if (i < tot) {
        const int r = i / nd, c = i - r * nd;
        dst[u] = r * lds + c;
        v[u] = x[(long long)min(w0 + r, W - 1) * nd + c];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * PREP_THREADS < tot) sx[dst[u]] = v[u];
  }
  __syncthreads();
  if (threadIdx.x >= nw) return;
  const double* xr = sx + threadIdx.x * lds;
  const int w = w0 + threadIdx.x;

  prepare_walker(p, tabs, xr, w, W, coef, ldc, prior);
}

// lnl = like + prior with like = sum scalar coefs (+ chi2/2), cosmoslik.py:141; +inf on a hard prior (:135-137)
__global__ void k_combine(csl_program_t p, int W, const double* __restrict__ coef, int ldc,
                          const double* __restrict__ prior, const double* __restrict__ chi2,
                          double* __restrict__ lnl) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  lnl[w] = combine_walker(p, w, coef, ldc, prior[w], chi2 != nullptr, chi2 != nullptr ? chi2[w] : 0.0);
}

// the same with chi^2 still in its per-32-row partial sums (explicit-inverse mode): they are added here, in the order
// k_sum_partials adds them, so the separate 5 us reduction launch of the staged / end-to-end path disappears
__global__ void k_combine_parts(csl_program_t p, int W, const double* __restrict__ coef, int ldc,
                                const double* __restrict__ prior, const double* __restrict__ part, int nblk, int ldp,
                                double* __restrict__ lnl) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double s = 0.0;
  for (int i = 0; i < nblk; ++i) s += part[(long long)i * ldp + w];
  lnl[w] = combine_walker(p, w, coef, ldc, prior[w], true, s);
}

// resid_mode 2: R[i, w] = coef[direct_coef[i], w] - data[i]; rows n..n_pad and columns W..Wp are zero
__global__ void k_direct_residual(csl_program_t p, int W, int Wp, const double* __restrict__ coef, int ldc,
                                  double* __restrict__ R, int ldr) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (w >= Wp) return;
  double v = 0.0;
  if (i < p.n && w < W) v = __dsub_rn(coef[(long long)p.direct_coef[i] * ldc + w], p.data[i]);
  R[(long long)i * ldr + w] = v;
}

// W walkers are evaluated; coefficient columns W..Wp-1 (tile padding) replay walker W-1; ldc is the stride
int csl_eval_prepare_ld(const csl_program_t* p, int W, int Wp, const double* x, double* coef, int ldc, double* prior,
                        void* stream) {
  CSL_REQUIRE(p && x && prior && W > 0, "csl_eval_prepare: null argument or W <= 0");
  CSL_REQUIRE(p->ndim >= 1 && p->ndim <= CSL_MAX_DIM, "csl_eval_prepare: ndim out of range");
  CSL_REQUIRE(p->ncoef == 0 || (coef && ldc >= Wp && Wp >= W), "csl_eval_prepare: coef buffer / ldc");
  int grid = (Wp + PREP_THREADS - 1) / PREP_THREADS;
  size_t smem = (size_t)PREP_THREADS * (p->ndim + 1) * sizeof(double);
  const int stage_tables = csl_small_stage_bytes(p) > 0 ? 1 : 0;
  smem += csl_small_stage_bytes(p);
  // the x tile of 128 walkers passes the 48 KB default limit of dynamic shared memory from ndim = 48 on
  // (66.5 KB at CSL_MAX_DIM = 64): opt in to what this program needs (per device; cheap, so every time)
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  k_prepare<<<grid, PREP_THREADS, smem, (cudaStream_t)stream>>>(*p, W, Wp, x, coef, ldc, prior, stage_tables);
  return csl_check_launch("k_prepare");
}

extern "C" int csl_eval_prepare(const csl_program_t* p, int W, const double* x, double* coef, int ldc,
                                double* prior, void* stream) {
  CSL_REQUIRE(p && (p->ncoef == 0 || ldc >= W), "csl_eval_prepare: coef buffer / ldc");
  return csl_eval_prepare_ld(p, W, (p->ncoef > 0) ? ldc : W, x, coef, ldc, prior, stream);   // columns W..ldc-1 replay walker W-1
}

extern "C" int csl_combine(const csl_program_t* p, int W, const double* coef, int ldc, const double* prior,
                           const double* chi2, double* lnl, void* stream) {
  CSL_REQUIRE(p && prior && lnl && W > 0, "csl_combine: null argument");
  k_combine<<<(W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, W, coef, ldc, prior, chi2, lnl);
  return csl_check_launch("k_combine");
}

extern "C" int csl_combine_parts(const csl_program_t* p, int W, const double* coef, int ldc, const double* prior,
                                 const double* part, int ldp, double* lnl, void* stream) {
  CSL_REQUIRE(p && prior && lnl && part && W > 0 && ldp >= W && p->n_pad > 0, "csl_combine_parts: bad argument");
  k_combine_parts<<<(W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p, W, coef, ldc, prior, part, p->n_pad / 32, ldp, lnl);
  return csl_check_launch("k_combine_parts");
}

int csl_direct_residual_ld(const csl_program_t* p, int W, int Wp, const double* coef, int ldc, double* R, int ldr,
                           void* stream) {
  CSL_REQUIRE(p && coef && R && W > 0 && ldr >= Wp && Wp >= W, "csl_direct_residual: bad argument");
  CSL_REQUIRE(p->resid_mode == 2, "csl_direct_residual: program has no direct residual");
  dim3 grid((Wp + 255) / 256, p->n_pad);
  k_direct_residual<<<grid, 256, 0, (cudaStream_t)stream>>>(*p, W, Wp, coef, ldc, R, ldr);
  return csl_check_launch("k_direct_residual");
}

extern "C" int csl_direct_residual(const csl_program_t* p, int W, const double* coef, int ldc, double* R,
                                   int ldr, void* stream) {
  return csl_direct_residual_ld(p, W, ldr, coef, ldc, R, ldr, stream);
}
