// Native Metropolis-Hastings step loop for ANY compiled posterior, and the on-device proposal adaptation.
//   reference: cosmoslik_plugins/samplers/metropolis_hastings.py:141-164 (draw_x, _mcmc),
//              :188-241 (master/worker block protocol), :249-253 (get_new_cov)
// One MH step of all chains on this GPU is
//     k_mh_propose_prepare  (draw_x + priors + coefficient bytecode; thread per chain)
//     residual kernel(s)    (binning.cu / prepare.cu, by resid_mode)
//     chi^2 kernel          (chi2.cu / covchi2.cu)
//     k_mh_combine_accept   (partial sums + lnl combine + accept rule + weights + row emission)
// enqueued from C with no Python in between (csl_mh_run), or captured once per block length into a CUDA graph
// that is relaunched for every block (csl_mh_graph_*): the step number lives in DEVICE memory (state->step), so
// the kernel arguments of a block never change.
// Arithmetic that decides chain contents is bit for bit that of the staged kernels k_mh_propose / k_mh_accept
// (same Philox counters, same fma order in z.F, same __dadd_rn / __dmul_rn / __dsub_rn in the accept rule).
#include "eval_device.cuh"
#include "philox.cuh"
#include <stdlib.h>

#define MHP_THREADS 64          // chains per CTA of the proposal kernel
#define MHA_THREADS 128         // chains per CTA of the accept kernel

__host__ __device__ constexpr int mh_row_stride(int nd) { return nd | 1; }   // odd: conflict-free 8-byte rows

// ---------------------------------------------------------------------------------------------
// draw_x (:141-142) + priors + bytecode
//   mode 0: delta injected;  1: standard normals injected;  2: Philox normals
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MHP_THREADS)
k_mh_propose_prepare(csl_program_t p, int W, int Wp, const double* __restrict__ cur_x, double* __restrict__ test_x,
                     const double* __restrict__ F, int mode, const double* __restrict__ inj, uint64_t seed,
                     const uint32_t* __restrict__ step_ptr, uint32_t step_off, uint32_t chain0,
                     double* __restrict__ coef, int ldc, double* __restrict__ prior, int stage_tables) {
  extern __shared__ __align__(16) double sm[];
  const int nd = p.ndim, lds = mh_row_stride(nd);
  double* sF = sm;                               // [nd*nd]
  double* zt = sF + nd * nd;                     // [MHP_THREADS][lds]  normals (or offsets)
  double* xt = zt + MHP_THREADS * lds;           // [MHP_THREADS][lds]  cur_x, then test_x
  // bytecode + prior tables: staged with the same round of loads as the proposal factor and the chain states
  const csl_small_tables tabs = stage_tables ? stage_small_tables(p, xt + MHP_THREADS * lds, threadIdx.x, MHP_THREADS)
                                             : small_tables_global(p);
  const int w0 = blockIdx.x * MHP_THREADS;
  const int nw = min(MHP_THREADS, Wp - w0);
  const uint32_t step = *step_ptr + step_off;
  if (mode != 0)
    for (int i = threadIdx.x; i < nd * nd; i += MHP_THREADS) sF[i] = F[i];
  // coalesced staging; columns W..Wp-1 (tile padding) replay chain W-1
  const int tot = nw * nd;
  for (int i0 = threadIdx.x; i0 < tot; i0 += 4 * MHP_THREADS) {      // four independent loads in flight per trip
    double vx[4], vz[4];
    int dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * MHP_THREADS;
      if (i < tot) {
        const int r = i / nd, c = i - r * nd;
        const long long g = (long long)min(w0 + r, W - 1) * nd + c;
        dst[u] = r * lds + c;
        vx[u] = cur_x[g];
        if (mode != 2) vz[u] = inj[(long long)step_off * W * nd + g];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u * MHP_THREADS < tot) {
        xt[dst[u]] = vx[u];
        if (mode != 2) zt[dst[u]] = vz[u];
      }
  }
  __syncthreads();
  const int tl = threadIdx.x, w = w0 + tl;
  if (tl < nw) {
    double* z = zt + tl * lds;
    double* x = xt + tl * lds;
    if (mode == 2) {
      const uint32_t chain = chain0 + (uint32_t)min(w, W - 1);
      for (int pr = 0; 2 * pr < nd; ++pr) {
        double z0, z1;
        philox_normal_pair(seed, chain, (uint32_t)pr, step, z0, z1);
        z[2 * pr] = z0;
        if (2 * pr + 1 < nd) z[2 * pr + 1] = z1;
      }
    }
    if (mode == 0) {
      for (int i = 0; i < nd; ++i) x[i] = __dadd_rn(x[i], z[i]);
    } else {
      // d_i = sum_k z_k F[k][i], k ascending from d = 0: the order of k_mh_propose
      for (int i = 0; i < nd; ++i) {
        double d = 0.0;
        for (int k = 0; k < nd; ++k) d = fma(z[k], sF[k * nd + i], d);
        x[i] = __dadd_rn(x[i], d);
      }
    }
    prepare_walker(p, tabs, x, w, W, coef, ldc, prior);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nw * nd; i += MHP_THREADS) {
    const int r = i / nd, c = i - r * nd;
    if (w0 + r < W) test_x[(long long)(w0 + r) * nd + c] = xt[r * lds + c];
  }
}

// ---------------------------------------------------------------------------------------------
// lnl(test) + accept / reject + weights + rows  (:154-164)
// ---------------------------------------------------------------------------------------------
struct csl_mh_combine {
  csl_program_t p;
  const double* coef; int ldc;
  const double* prior; const double* chi2; const double* part; int nblk, ldp;
};

__global__ void __launch_bounds__(MHA_THREADS)
k_mh_combine_accept(int W, int nd, double* __restrict__ cur_x, double* __restrict__ cur_lnl,
                    double* __restrict__ cur_weight, const double* __restrict__ test_x, double* __restrict__ test_lnl,
                    const double* __restrict__ uin, uint64_t seed, const uint32_t* __restrict__ step_ptr,
                    uint32_t step_off, uint32_t chain0, double temp, double max_weight, int yield_rejected,
                    double* __restrict__ rows, int32_t* __restrict__ nrows, int cap, uint8_t* __restrict__ accepted,
                    double* __restrict__ margin, unsigned long long* __restrict__ nacc, csl_mh_combine fc) {
  // per chain of the CTA: bit 0 accepted, bit 1 the current state is emitted, bit 2 the rejected point is emitted
  __shared__ unsigned char sflag[MHA_THREADS];
  __shared__ int sslot[MHA_THREADS];             // slot of the first row this chain emits in this step
  const int k0 = blockIdx.x * MHA_THREADS, w = k0 + threadIdx.x;
  const uint32_t step = *step_ptr + step_off;
  const int rl = 2 + nd;
  unsigned flag = 0;
  bool acc = false;
  if (w < W) {
    double chi2 = 0.0;
    if (fc.part) for (int i = 0; i < fc.nblk; ++i) chi2 += fc.part[(long long)i * fc.ldp + w];
    else if (fc.chi2) chi2 = fc.chi2[w];
    const double tl = combine_walker(fc.p, w, fc.coef, fc.ldc, fc.prior[w], fc.part != nullptr || fc.chi2 != nullptr, chi2);
    test_lnl[w] = tl;
    const double cl = cur_lnl[w];
    double cw = cur_weight[w];
    const double uu = uin ? uin[(long long)step_off * W + w] : philox_mh_uniform(seed, chain0 + (uint32_t)w, step);
    const double lhs = log(uu), rhs = __dmul_rn(temp, __dsub_rn(cl, tl));
    acc = lhs < rhs;                 // inf - inf = nan -> false (reject), as in the reference
    int slot = nrows[w], n_emit = 0;
    sslot[threadIdx.x] = slot;
    if (acc) {
      if (cl != CUDART_INF) {        // yield(sample(cur_x, cur_lnl, cur_weight))
        flag |= 2u;
        if (slot < cap) { double* r = rows + ((long long)w * cap + slot) * rl; r[0] = cl; r[1] = cw; }
        ++n_emit;
      }
      cur_lnl[w] = tl;
      cur_weight[w] = 1.0;
      flag |= 1u;
    } else {
      if (yield_rejected) {          // yield(sample(test_x, test_lnl, 0))  (:160)
        flag |= 4u;
        if (slot < cap) { double* r = rows + ((long long)w * cap + slot) * rl; r[0] = tl; r[1] = 0.0; }
        ++n_emit;
      }
      cw += 1.0;
      if (cw >= max_weight) {        // yield(sample(cur_x, cur_lnl, cur_weight)); cur_weight = 0
        flag |= 2u;
        if (slot + n_emit < cap) { double* r = rows + ((long long)w * cap + slot + n_emit) * rl; r[0] = cl; r[1] = cw; }
        ++n_emit;
        cw = 0.0;
      }
      cur_weight[w] = cw;
    }
    if (n_emit) nrows[w] = slot + n_emit;      // counts past cap signal overflow to the host
    if (accepted) accepted[w] = acc ? 1 : 0;
    if (margin) margin[w] = lhs - rhs;
  }
  sflag[threadIdx.x] = (unsigned char)flag;
  const int c = __syncthreads_count(acc);
  if (nacc && threadIdx.x == 0 && c) atomicAdd(nacc, (unsigned long long)c);
  // the CTA's chains as one contiguous range of positions: emitted rows get their x, accepted chains move
  const int nw = min(MHA_THREADS, W - k0);
  const long long base = (long long)k0 * nd;
  for (int idx = threadIdx.x; idx < nw * nd; idx += MHA_THREADS) {
    const int r = idx / nd, cidx = idx - r * nd;
    const unsigned f = sflag[r];
    if (!f) continue;
    const double xo = cur_x[base + idx], xn = test_x[base + idx];
    int slot = sslot[r];
    const long long rb = (long long)(k0 + r) * cap;
    if (f & 4u) { if (slot < cap) rows[(rb + slot) * rl + 2 + cidx] = xn; ++slot; }
    if (f & 2u) { if (slot < cap) rows[(rb + slot) * rl + 2 + cidx] = xo; }
    if (f & 1u) cur_x[base + idx] = xn;
  }
}

__global__ void k_step_advance(uint32_t* step, uint32_t n) { *step += n; }

// ---------------------------------------------------------------------------------------------
// proposal adaptation on the device (:249-253 + numpy's factorisation inside draw_x, :141-142)
// ---------------------------------------------------------------------------------------------
// ONE pass over the last half of every chain's rows:
//   partial[b] = [ sum w,  sum w (x - ref),  sum w (x - ref)(x - ref)^T ]      (1 + nd + nd^2 doubles)
// `ref` is any fixed point near the chains (the start point): the shift keeps the raw second moments well
// conditioned, and being the SAME on every rank it lets one all-reduce of the sums replace the two-pass
// (mean first) scheme.  Reduction order is fixed (bitwise reproducible for a given sharding).
#define FM_BLOCKS 592
#define FM_TILE 64
#define FM_THREADS 256
__global__ void __launch_bounds__(FM_THREADS)
k_mh_moments_fused(int W, int nd, const double* __restrict__ rows, const int32_t* __restrict__ nrows, int cap,
                   const double* __restrict__ ref, double* __restrict__ partial) {
  extern __shared__ double sx[];            // [FM_TILE][2 + nd]: (lnl, w, x - ref ...)
  const int m = 1 + nd + nd * nd, rl = 2 + nd;
  constexpr int NACC = (CSL_MAX_DIM * CSL_MAX_DIM + FM_THREADS - 1) / FM_THREADS;
  double acc[NACC];
  int pi[NACC], pk[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    acc[q] = 0.0;
    const int e = threadIdx.x + q * FM_THREADS;
    pi[q] = (e < nd * nd) ? e / nd : -1;
    pk[q] = (e < nd * nd) ? e - (e / nd) * nd : 0;
  }
  double a1 = 0.0;                          // thread 0: sum w; thread i in 1..nd: sum w (x_i - ref_i)
  for (int w = blockIdx.x; w < W; w += FM_BLOCKS) {
    const int nr = min(nrows[w], cap);
    for (int t0 = nr / 2; t0 < nr; t0 += FM_TILE) {
      const int nt = min(FM_TILE, nr - t0);
      const double* src = rows + ((long long)w * cap + t0) * rl;
      __syncthreads();
      for (int j = threadIdx.x; j < nt * rl; j += FM_THREADS) {
        const int c = j % rl;
        sx[j] = c >= 2 ? src[j] - ref[c - 2] : src[j];
      }
      __syncthreads();
      if (threadIdx.x <= nd) {
        for (int r = 0; r < nt; ++r) {
          const double wt = sx[r * rl + 1];
          a1 = (threadIdx.x == 0) ? a1 + wt : fma(wt, sx[r * rl + 1 + threadIdx.x], a1);
        }
      }
#pragma unroll
      for (int q = 0; q < NACC; ++q) {
        if (pi[q] >= 0) {
          double s_ = acc[q];
          for (int r = 0; r < nt; ++r) s_ = fma(sx[r * rl + 1] * sx[r * rl + 2 + pi[q]], sx[r * rl + 2 + pk[q]], s_);
          acc[q] = s_;
        }
      }
    }
  }
  double* out = partial + (long long)blockIdx.x * m;
  if (threadIdx.x <= nd) out[threadIdx.x] = a1;
#pragma unroll
  for (int q = 0; q < NACC; ++q)
    if (pi[q] >= 0) out[1 + nd + threadIdx.x + q * FM_THREADS] = acc[q];
}

__global__ void k_fm_reduce(int m, const double* __restrict__ partial, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m) return;
  double s = 0.0;
  for (int b = 0; b < FM_BLOCKS; ++b) s += partial[(long long)b * m + e];
  out[e] = s;
}

// mom = [S0, S1 (nd), S2 (nd x nd)] about `ref` (already summed over ranks)
//   cov  = (S2 - S1 S1^T / S0) / (S0 - 1)                        (chains.py:507-512)
//   F    : F^T F = cov / nd * scale^2   -- F = L^T with L the lower Cholesky factor, so that
//          test_x = cur_x + z F has the covariance multivariate_normal(cur_x, cov_est / ndim * scale^2) has.
//          (numpy factors by SVD; ANY factor gives the same proposal distribution.  The host SVD factor is
//          kept for runs with injected normal streams, where the proposals themselves must match the reference.)
// A non-positive pivot (a direction no chain has moved in) zeroes its column instead of producing NaN.
// One CTA; cov (symmetric) and F are written as [nd, nd] row-major.
__global__ void __launch_bounds__(256)
k_mh_adapt_factor(int nd, const double* __restrict__ mom, double scale2_over_nd, double* __restrict__ cov_out,
                  double* __restrict__ F_out, double* __restrict__ mean_out, const double* __restrict__ ref) {
  __shared__ double A[CSL_MAX_DIM * (CSL_MAX_DIM + 1)];
  const int ld = nd + 1;
  const double s0 = mom[0];
  for (int e = threadIdx.x; e < nd * nd; e += blockDim.x) {
    const int i = e / nd, k = e - i * nd;
    const double c = (mom[1 + nd + e] - mom[1 + i] * mom[1 + k] / s0) / (s0 - 1.0);
    A[i * ld + k] = c;
  }
  __syncthreads();
  // symmetrise (the two triangles were summed in different orders) and publish the covariance
  for (int e = threadIdx.x; e < nd * nd; e += blockDim.x) {
    const int i = e / nd, k = e - i * nd;
    cov_out[e] = 0.5 * (A[i * ld + k] + A[k * ld + i]);
  }
  if (mean_out)
    for (int i = threadIdx.x; i < nd; i += blockDim.x) mean_out[i] = ref[i] + mom[1 + i] / s0;
  __syncthreads();
  for (int e = threadIdx.x; e < nd * nd; e += blockDim.x) {
    const int i = e / nd, k = e - i * nd;
    A[i * ld + k] = cov_out[e] * scale2_over_nd;
  }
  __syncthreads();
  // right-looking Cholesky in shared memory (lower triangle)
  double dmax = 0.0;
  for (int i = 0; i < nd; ++i) dmax = fmax(dmax, A[i * ld + i]);
  const double tiny = dmax * 1e-12;
  for (int j = 0; j < nd; ++j) {
    const double piv = A[j * ld + j];
    const bool ok = piv > tiny;
    const double d = ok ? sqrt(piv) : 0.0, inv = ok ? 1.0 / d : 0.0;
    __syncthreads();
    for (int i = j + threadIdx.x; i < nd; i += blockDim.x) A[i * ld + j] = (i == j) ? d : A[i * ld + j] * inv;
    __syncthreads();
    const int rem = nd - j - 1;
    for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
      const int i = j + 1 + e / rem, k = j + 1 + e % rem;
      if (k <= i) A[i * ld + k] = fma(-A[i * ld + j], A[k * ld + j], A[i * ld + k]);
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < nd * nd; e += blockDim.x) {
    const int k = e / nd, i = e - k * nd;            // F[k][i] = L[i][k]
    F_out[e] = (i >= k) ? A[i * ld + k] : 0.0;
  }
}

// =============================================================================================
static inline int round_up_i(int v, int m) { return (v + m - 1) / m * m; }

static int mh_step(const csl_program_t* p, const csl_mh_t* m, int off, void* stream) {
  const int nd = m->ndim, W = m->W;
  const bool quad = p->resid_mode != 0;
  const int Wp = quad ? round_up_i(W, CSL_BN) : W;
  const int mode = m->inj_z ? (m->inj_is_delta ? 0 : 1) : 2;
  cudaStream_t s = (cudaStream_t)stream;
  const int lds = mh_row_stride(nd);
  size_t smem = (size_t)(nd * nd + 2 * MHP_THREADS * lds) * sizeof(double);
  const int stage_tables = csl_small_stage_bytes(p) > 0 ? 1 : 0;
  smem += csl_small_stage_bytes(p);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(k_mh_propose_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  const int ldc = Wp;
  k_mh_propose_prepare<<<(Wp + MHP_THREADS - 1) / MHP_THREADS, MHP_THREADS, smem, s>>>(
      *p, W, p->ncoef > 0 ? Wp : W, m->cur_x, m->test_x, m->F, mode, m->inj_z, m->seed, m->step, (uint32_t)off,
      m->chain0, m->ws_coef, ldc, m->ws_prior, stage_tables);
  int rc = csl_check_launch("k_mh_propose_prepare");
  if (rc) return rc;
  const double* chi2 = nullptr;
  const double* part = nullptr;
  if (quad) {
    if (p->resid_mode == 1 || p->resid_mode == 3) rc = csl_bin_residual(p, Wp, m->ws_coef, Wp, m->ws_R, Wp, stream);
    else rc = csl_direct_residual(p, W, m->ws_coef, Wp, m->ws_R, Wp, stream);
    if (rc) return rc;
    if (p->resid_mode == 3) {
      rc = csl_cov_chi2(p, W, m->ws_R, Wp, m->ws_chi2, stream);
      chi2 = m->ws_chi2;
    } else {
      rc = csl_chi2_impl(p, Wp, m->ws_R, Wp, m->ws_chi2, m->ws_part, false, stream);
      if (p->chi2_mode == 1) part = m->ws_part; else chi2 = m->ws_chi2;
    }
    if (rc) return rc;
  }
  csl_mh_combine fc;
  fc.p = *p; fc.coef = m->ws_coef; fc.ldc = ldc; fc.prior = m->ws_prior; fc.chi2 = chi2; fc.part = part;
  fc.nblk = p->n_pad / 32; fc.ldp = Wp;
  k_mh_combine_accept<<<(W + MHA_THREADS - 1) / MHA_THREADS, MHA_THREADS, 0, s>>>(
      W, nd, m->cur_x, m->cur_lnl, m->cur_weight, m->test_x, m->test_lnl, m->inj_u, m->seed, m->step, (uint32_t)off,
      m->chain0, m->temp, m->max_weight, m->yield_rejected, m->rows, m->nrows, m->cap, m->accepted, m->margin,
      reinterpret_cast<unsigned long long*>(m->naccepted), fc);
  return csl_check_launch("k_mh_combine_accept");
}

static int mh_validate(const csl_program_t* p, const csl_mh_t* m, int nsteps) {
  CSL_REQUIRE(p && m && nsteps >= 0, "csl_mh_run: null argument");
  CSL_REQUIRE(m->ndim == p->ndim && m->ndim >= 1 && m->ndim <= CSL_MAX_DIM && m->W > 0 && m->cap > 0,
              "csl_mh_run: bad state (ndim / W / cap)");
  CSL_REQUIRE(m->cur_x && m->cur_lnl && m->cur_weight && m->test_x && m->test_lnl && m->rows && m->nrows && m->step &&
                  m->ws_prior, "csl_mh_run: state buffer missing");
  CSL_REQUIRE((m->inj_z && m->inj_is_delta) || m->F, "csl_mh_run: proposal factor missing");
  CSL_REQUIRE(p->ncoef == 0 || m->ws_coef, "csl_mh_run: coefficient workspace missing");
  CSL_REQUIRE(p->resid_mode == 0 || (m->ws_R && m->ws_chi2), "csl_mh_run: workspace missing for the quadratic form");
  CSL_REQUIRE(p->resid_mode == 0 || p->resid_mode == 3 || p->chi2_mode == 0 || m->ws_part,
              "csl_mh_run: partial-sum workspace missing");
  return 0;
}

static int mh_enqueue(const csl_program_t* p, const csl_mh_t* m, int nsteps, void* stream) {
  for (int i = 0; i < nsteps; ++i) {
    int rc = mh_step(p, m, i, stream);
    if (rc) return rc;
  }
  if (nsteps > 0) {
    k_step_advance<<<1, 1, 0, (cudaStream_t)stream>>>(m->step, (uint32_t)nsteps);
    return csl_check_launch("k_step_advance");
  }
  return 0;
}

extern "C" int csl_sizeof_mh(void) { return (int)sizeof(csl_mh_t); }

extern "C" int csl_mh_run(const csl_program_t* p, const csl_mh_t* m, int nsteps, void* stream) {
  int rc = mh_validate(p, m, nsteps);
  if (rc) return rc;
  return mh_enqueue(p, m, nsteps, stream);
}

// ---- CUDA graph of one block of steps ------------------------------------------------------------------------
struct csl_graph {
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStream_t capture_stream;
  int nodes;
};

extern "C" int csl_mh_graph_create(const csl_program_t* p, const csl_mh_t* m, int nsteps, void** out) {
  CSL_REQUIRE(out, "csl_mh_graph_create: null output");
  *out = nullptr;
  int rc = mh_validate(p, m, nsteps);
  if (rc) return rc;
  CSL_REQUIRE(nsteps > 0, "csl_mh_graph_create: nsteps must be positive");
  csl_graph* g = (csl_graph*)calloc(1, sizeof(csl_graph));
  if (!g) return csl_fail(CSL_E_BADARG, "csl_mh_graph_create: out of host memory");
  cudaError_t e = cudaStreamCreateWithFlags(&g->capture_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { free(g); return csl_fail((int)e, cudaGetErrorString(e)); }
  // every kernel's shared-memory opt-in is set BEFORE capture by one eager dry launch of nothing: the launch
  // helpers set the attribute unconditionally when they need it, which is legal during capture
  e = cudaStreamBeginCapture(g->capture_stream, cudaStreamCaptureModeThreadLocal);
  if (e != cudaSuccess) { cudaStreamDestroy(g->capture_stream); free(g); return csl_fail((int)e, cudaGetErrorString(e)); }
  rc = mh_enqueue(p, m, nsteps, g->capture_stream);
  e = cudaStreamEndCapture(g->capture_stream, &g->graph);
  if (rc || e != cudaSuccess) {
    if (g->graph) cudaGraphDestroy(g->graph);
    cudaStreamDestroy(g->capture_stream);
    free(g);
    return rc ? rc : csl_fail((int)e, cudaGetErrorString(e));
  }
  size_t nn = 0;
  cudaGraphGetNodes(g->graph, nullptr, &nn);
  g->nodes = (int)nn;
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g->graph); cudaStreamDestroy(g->capture_stream); free(g);
    return csl_fail((int)e, cudaGetErrorString(e));
  }
  *out = g;
  return 0;
}

extern "C" int csl_graph_launch(void* graph, void* stream) {
  CSL_REQUIRE(graph, "csl_graph_launch: null graph");
  cudaError_t e = cudaGraphLaunch(((csl_graph*)graph)->exec, (cudaStream_t)stream);
  if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  return 0;
}

extern "C" int csl_graph_nodes(void* graph) { return graph ? ((csl_graph*)graph)->nodes : 0; }

extern "C" int csl_graph_destroy(void* graph) {
  if (!graph) return 0;
  csl_graph* g = (csl_graph*)graph;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  if (g->capture_stream) cudaStreamDestroy(g->capture_stream);
  free(g);
  return 0;
}

// ---- adaptation ------------------------------------------------------------------------------------------------
extern "C" int csl_mh_moments_fused(int W, int ndim, const double* rows, const int32_t* nrows, int cap,
                                    const double* ref, double* partial, double* out, void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && rows && nrows && ref && partial && out && cap > 0,
              "csl_mh_moments_fused: bad argument");
  static_assert((size_t)FM_TILE * (2 + CSL_MAX_DIM) * sizeof(double) <= 48 * 1024, "row tile fits the default dynamic shared memory");
  const int m = 1 + ndim + ndim * ndim;
  cudaStream_t s = (cudaStream_t)stream;
  k_mh_moments_fused<<<FM_BLOCKS, FM_THREADS, (size_t)FM_TILE * (2 + ndim) * sizeof(double), s>>>(W, ndim, rows, nrows, cap,
                                                                                                 ref, partial);
  k_fm_reduce<<<(m + 127) / 128, 128, 0, s>>>(m, partial, out);
  return csl_check_launch("k_mh_moments_fused");
}

extern "C" int csl_mh_adapt_factor(int ndim, const double* mom, double proposal_scale, const double* ref, double* cov_out,
                                   double* F_out, double* mean_out, void* stream) {
  CSL_REQUIRE(ndim >= 1 && ndim <= CSL_MAX_DIM && mom && cov_out && F_out && (mean_out == nullptr || ref),
              "csl_mh_adapt_factor: bad argument");
  k_mh_adapt_factor<<<1, 256, 0, (cudaStream_t)stream>>>(ndim, mom, proposal_scale * proposal_scale / ndim, cov_out, F_out,
                                                         mean_out, ref);
  return csl_check_launch("k_mh_adapt_factor");
}
