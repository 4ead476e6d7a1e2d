// Sampler kernels: batched Metropolis-Hastings and the affine-invariant stretch move.
//   reference: cosmoslik_plugins/samplers/metropolis_hastings.py:141-164, 249-253
//              cosmoslik_plugins/samplers/emcee.py:66-88 (+ third-party emcee 2.x _propose_stretch)
// One warp per walker: lane i owns parameter i (and i+32), so every global access to a walker's
// row is one coalesced 8*ndim-byte segment.  These kernels are HBM-bound: ~(24 ndim + 32) B per
// walker per step.
#include "eval_device.cuh"
#include "philox.cuh"
#include <stdlib.h>

#define SAMP_THREADS 256
#define SAMP_WARPS (SAMP_THREADS / 32)

__device__ __forceinline__ int warp_global() { return (blockIdx.x * SAMP_THREADS + threadIdx.x) >> 5; }

// ---------------------------------------------------------------------------------------------
// draw_x  (metropolis_hastings.py:141-142)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SAMP_THREADS)
k_mh_propose(int W, int nd, const double* __restrict__ cur_x, double* __restrict__ test_x, int mode,
             const double* __restrict__ inj, const double* __restrict__ F, uint64_t seed, uint32_t step,
             uint32_t chain0) {
  extern __shared__ double sF[];
  if (mode != 0)
    for (int i = threadIdx.x; i < nd * nd; i += SAMP_THREADS) sF[i] = F[i];
  __syncthreads();
  const int w = warp_global(), lane = threadIdx.x & 31;
  if (w >= W) return;
  const long long base = (long long)w * nd;
  const int i0 = lane, i1 = lane + 32;
  double x0 = (i0 < nd) ? cur_x[base + i0] : 0.0, x1 = (i1 < nd) ? cur_x[base + i1] : 0.0;
  double d0 = 0.0, d1 = 0.0;
  if (mode == 0) {
    if (i0 < nd) d0 = inj[base + i0];
    if (i1 < nd) d1 = inj[base + i1];
  } else {
    double za = 0.0, zb = 0.0;  // mode 1: z[lane], z[lane+32]; mode 2: pair (z[2 lane], z[2 lane + 1])
    if (mode == 1) {
      if (i0 < nd) za = inj[base + i0];
      if (i1 < nd) zb = inj[base + i1];
    } else if (2 * lane < nd) {
      philox_normal_pair(seed, chain0 + (uint32_t)w, (uint32_t)lane, step, za, zb);
    }
    for (int k = 0; k < nd; ++k) {
      double zk = (mode == 1) ? shfl_d(k < 32 ? za : zb, k & 31) : shfl_d((k & 1) ? zb : za, k >> 1);
      if (i0 < nd) d0 = fma(zk, sF[k * nd + i0], d0);
      if (i1 < nd) d1 = fma(zk, sF[k * nd + i1], d1);
    }
  }
  if (i0 < nd) test_x[base + i0] = __dadd_rn(x0, d0);
  if (i1 < nd) test_x[base + i1] = __dadd_rn(x1, d1);
}

// row emission shared by the MH and ensemble kernels: (lnl, weight, x...) at slot nrows[w]
__device__ __forceinline__ void emit_row(double* rows, int32_t* nrows, int cap, int w, int nd, int lane,
                                         double lnl, double weight, double x0, double x1) {
  int slot = nrows[w];
  __syncwarp();
  if (slot < cap) {
    double* r = rows + ((long long)w * cap + slot) * (2 + nd);
    if (lane == 0) { r[0] = lnl; r[1] = weight; }
    if (lane < nd) r[2 + lane] = x0;
    if (lane + 32 < nd) r[2 + lane + 32] = x1;
  }
  if (lane == 0) nrows[w] = slot + 1;   // counts past cap signal overflow to the host
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// accept / reject + weight bookkeeping  (metropolis_hastings.py:154-164)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SAMP_THREADS)
k_mh_accept(int W, int nd, double* __restrict__ cur_x, double* __restrict__ cur_lnl,
            double* __restrict__ cur_weight, const double* __restrict__ test_x,
            const double* __restrict__ test_lnl, const double* __restrict__ u, uint64_t seed, uint32_t step,
            uint32_t chain0, double temp, double max_weight, double* __restrict__ rows,
            int32_t* __restrict__ nrows, int cap, uint8_t* __restrict__ accepted, double* __restrict__ margin) {
  const int w = warp_global(), lane = threadIdx.x & 31;
  if (w >= W) return;
  const long long base = (long long)w * nd;
  const double cl = cur_lnl[w], tl = test_lnl[w];
  double cw = cur_weight[w];
  const double uu = u ? u[w] : philox_mh_uniform(seed, chain0 + (uint32_t)w, step);
  const double lhs = log(uu), rhs = __dmul_rn(temp, __dsub_rn(cl, tl));
  const bool acc = lhs < rhs;   // inf - inf = nan -> false (reject), as in the reference
  double x0 = (lane < nd) ? cur_x[base + lane] : 0.0, x1 = (lane + 32 < nd) ? cur_x[base + lane + 32] : 0.0;
  if (acc) {
    if (cl != CUDART_INF) emit_row(rows, nrows, cap, w, nd, lane, cl, cw, x0, x1);
    if (lane < nd) cur_x[base + lane] = test_x[base + lane];
    if (lane + 32 < nd) cur_x[base + lane + 32] = test_x[base + lane + 32];
    if (lane == 0) { cur_lnl[w] = tl; cur_weight[w] = 1.0; }
  } else {
    cw += 1.0;
    if (cw >= max_weight) {
      emit_row(rows, nrows, cap, w, nd, lane, cl, cw, x0, x1);
      cw = 0.0;
    }
    if (lane == 0) cur_weight[w] = cw;
  }
  if (lane == 0) {
    if (accepted) accepted[w] = acc ? 1 : 0;
    if (margin) margin[w] = lhs - rhs;
  }
}

// ---------------------------------------------------------------------------------------------
// whole MH run for a dense Gaussian posterior: nsteps steps in one launch, one warp per chain
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SAMP_THREADS)
k_mh_gauss_run(csl_program_t p, int W, int nsteps, const double* __restrict__ Lfac,
               const double* __restrict__ mean, const int32_t* __restrict__ perm, double* __restrict__ cur_x,
               double* __restrict__ cur_lnl, double* __restrict__ cur_weight, const double* __restrict__ zin,
               int z_is_delta, const double* __restrict__ uin, const double* __restrict__ F, uint64_t seed,
               uint32_t step0, uint32_t chain0, double temp, double max_weight, double* __restrict__ rows,
               int32_t* __restrict__ nrows, int cap) {
  extern __shared__ double sm[];
  const int nd = p.ndim, n = p.n;
  double* sF = sm;                 // [nd*nd]
  double* sL = sF + nd * nd;       // [n*n]   lower Cholesky factor of the target covariance
  double* sInv = sL + n * n;       // [n]     1 / L_kk
  for (int i = threadIdx.x; i < nd * nd; i += SAMP_THREADS) sF[i] = F ? F[i] : 0.0;
  for (int i = threadIdx.x; i < n * n; i += SAMP_THREADS) sL[i] = Lfac[i];
  for (int i = threadIdx.x; i < n; i += SAMP_THREADS) sInv[i] = 1.0 / Lfac[i * n + i];
  __syncthreads();
  const int w = warp_global(), lane = threadIdx.x & 31;
  if (w >= W) return;
  const long long base = (long long)w * nd;
  double x = (lane < nd) ? cur_x[base + lane] : 0.0;
  double cl = cur_lnl[w], cw = cur_weight[w];
  const int myperm = (lane < n) ? perm[lane] : 0;
  const double mymean = (lane < n) ? mean[lane] : 0.0;

  for (int s = 0; s < nsteps; ++s) {
    const uint32_t step = step0 + (uint32_t)s;
    // ---- draw_x
    double d = 0.0;
    if (zin && z_is_delta) {
      if (lane < nd) d = zin[((long long)s * W + w) * nd + lane];
    } else {
      double za = 0.0, zb = 0.0;
      if (zin) { if (lane < nd) za = zin[((long long)s * W + w) * nd + lane]; }
      else if (2 * lane < nd) philox_normal_pair(seed, chain0 + (uint32_t)w, (uint32_t)lane, step, za, zb);
      for (int k = 0; k < nd; ++k) {
        double zk = zin ? shfl_d(za, k) : shfl_d((k & 1) ? zb : za, k >> 1);
        if (lane < nd) d = fma(zk, sF[k * nd + lane], d);
      }
    }
    const double tx = __dadd_rn(x, d);
    // ---- priors (likelihoods/priors.py:24-30), evaluated redundantly by every lane
    bool out = false;
    for (int b = 0; b < p.nbox; ++b) {
      double v = shfl_d(tx, p.box_idx[b]);
      if (!(p.box_lo[b] <= v && v <= p.box_hi[b])) out = true;
    }
    double pr = 0.0;
    for (int k = 0; k < p.ngauss; ++k) {
      double dv = __dsub_rn(shfl_d(tx, p.gauss_idx[k]), p.gauss_c[k]);
      double wv = p.gauss_w[k];
      pr = __dadd_rn(pr, __ddiv_rn(__ddiv_rn(__dmul_rn(dv, dv), 2.0), __dmul_rn(wv, wv)));
    }
    // ---- (x-mu)^T C^-1 (x-mu) = || L^-1 (x-mu) ||^2 : column-oriented forward substitution
    double r = shfl_d(tx, myperm) - mymean;
    double ss = 0.0;
    for (int k = 0; k < n; ++k) {
      double yk = shfl_d(r, k) * sInv[k];
      if (lane > k && lane < n) r = fma(-sL[lane * n + k], yk, r);
      ss = fma(yk, yk, ss);
    }
    const double tl = out ? CUDART_INF : __dadd_rn(0.5 * ss, pr);
    // ---- accept (metropolis_hastings.py:154-164)
    const double uu = uin ? uin[(long long)s * W + w] : philox_mh_uniform(seed, chain0 + (uint32_t)w, step);
    const bool acc = log(uu) < __dmul_rn(temp, __dsub_rn(cl, tl));
    if (acc) {
      if (cl != CUDART_INF) emit_row(rows, nrows, cap, w, nd, lane, cl, cw, x, 0.0);
      x = tx; cl = tl; cw = 1.0;
    } else {
      cw += 1.0;
      if (cw >= max_weight) {
        emit_row(rows, nrows, cap, w, nd, lane, cl, cw, x, 0.0);
        cw = 0.0;
      }
    }
  }
  if (lane < nd) cur_x[base + lane] = x;
  if (lane == 0) { cur_lnl[w] = cl; cur_weight[w] = cw; }
}

// ---------------------------------------------------------------------------------------------
// get_new_cov sufficient statistics (metropolis_hastings.py:249-253, chains.py:507-512)
// ---------------------------------------------------------------------------------------------
#define MOM_BLOCKS 592      // 4 CTAs per SM on 148 SMs
#define MOM_TILE 64         // rows staged per barrier pair
// pass 1 (mean == nullptr): partial[b][0] = sum w, [1..nd] = sum w x.
// pass 2: partial[b][1+nd + i*nd + k] = sum w (x_i - mean_i)(x_k - mean_k).
// The last half of a chain's rows is contiguous in memory: tiles of MOM_TILE rows are staged in shared
// memory with coalesced loads, then every thread owns a few (i,k) pairs and sweeps the tile.
__global__ void __launch_bounds__(SAMP_THREADS)
k_mh_moments(int W, int nd, const double* __restrict__ rows, const int32_t* __restrict__ nrows, int cap,
             const double* __restrict__ mean, double* __restrict__ partial, int burn) {
  extern __shared__ double sx[];            // [MOM_TILE][2 + nd]
  const int m = 1 + nd + nd * nd;
  const int rl = 2 + nd;
  constexpr int NACC = (CSL_MAX_DIM * CSL_MAX_DIM + SAMP_THREADS - 1) / SAMP_THREADS;
  double acc[NACC];
  int pi[NACC], pk[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    acc[q] = 0.0;
    int e = threadIdx.x + q * SAMP_THREADS;
    pi[q] = (e < nd * nd) ? e / nd : -1;
    pk[q] = (e < nd * nd) ? e - (e / nd) * nd : 0;
  }
  double mi[NACC], mk[NACC];
#pragma unroll
  for (int q = 0; q < NACC; ++q) {
    mi[q] = (mean && pi[q] >= 0) ? mean[pi[q]] : 0.0;
    mk[q] = (mean && pi[q] >= 0) ? mean[pk[q]] : 0.0;
  }
  double a1 = 0.0;  // pass 1: thread i <= nd accumulates element i
  for (int w = blockIdx.x; w < W; w += MOM_BLOCKS) {
    const int nr = min(nrows[w], cap);
    // burn < 0: the last half of the chain's rows (get_new_cov); burn >= 0: rows from `burn` on
    for (int t0 = (burn < 0 ? nr / 2 : min(burn, nr)); t0 < nr; t0 += MOM_TILE) {
      const int nt = min(MOM_TILE, nr - t0);
      const double* src = rows + ((long long)w * cap + t0) * rl;
      __syncthreads();
      for (int j = threadIdx.x; j < nt * rl; j += SAMP_THREADS) sx[j] = src[j];
      __syncthreads();
      if (mean == nullptr) {
        if (threadIdx.x <= nd) {
          for (int r = 0; r < nt; ++r) {
            const double wt = sx[r * rl + 1];
            a1 = (threadIdx.x == 0) ? a1 + wt : fma(wt, sx[r * rl + 1 + threadIdx.x], a1);
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
          if (pi[q] >= 0) {
            double s_ = acc[q];
            for (int r = 0; r < nt; ++r)
              s_ = fma(sx[r * rl + 1] * (sx[r * rl + 2 + pi[q]] - mi[q]), sx[r * rl + 2 + pk[q]] - mk[q], s_);
            acc[q] = s_;
          }
        }
      }
    }
  }
  double* out = partial + (long long)blockIdx.x * m;
  if (mean == nullptr) {
    if (threadIdx.x <= nd) out[threadIdx.x] = a1;
  } else {
#pragma unroll
    for (int q = 0; q < NACC; ++q)
      if (pi[q] >= 0) out[1 + nd + threadIdx.x + q * SAMP_THREADS] = acc[q];
  }
}

__global__ void k_reduce_partials(int m, int lo, int hi, const double* __restrict__ partial,
                                  double* __restrict__ out) {
  int e = lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= hi) return;
  double s = 0.0;
  for (int b = 0; b < MOM_BLOCKS; ++b) s += partial[(long long)b * m + e];
  out[e] = s;
}

// ---------------------------------------------------------------------------------------------
// stretch move (emcee 2.x _propose_stretch; Goodman & Weare 2010)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SAMP_THREADS)
k_stretch_propose(int Ns, int Nc, int nd, const double* __restrict__ s, const double* __restrict__ comp, double a,
                  const double* __restrict__ u_z, const int64_t* __restrict__ jin, uint64_t seed, uint32_t step,
                  uint32_t half, uint32_t walker0, double* __restrict__ q, double* __restrict__ zz) {
  const int k = warp_global(), lane = threadIdx.x & 31;
  if (k >= Ns) return;
  double uz, ua;
  int64_t j;
  if (u_z) { uz = u_z[k]; j = jin[k]; }
  else philox_stretch(seed, walker0 + (uint32_t)k, half, step, (uint32_t)Nc, uz, j, ua);
  // zz = ((a - 1.) * u + 1) ** 2. / a
  double tq = __dadd_rn(__dmul_rn(a - 1.0, uz), 1.0);
  const double z = __ddiv_rn(__dmul_rn(tq, tq), a);
  const long long sb = (long long)k * nd, cb = (long long)j * nd;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    if (i < nd) {
      double c = comp[cb + i];
      // q = c - zz * (c - s)
      q[sb + i] = __dsub_rn(c, __dmul_rn(z, __dsub_rn(c, s[sb + i])));
    }
  }
  if (lane == 0) zz[k] = z;
}

// Proposal + priors + coefficient bytecode in one launch (the native step loop, csl_stretch_run): one
// thread per walker computes its proposal row (same arithmetic as k_stretch_propose), keeps it in shared
// memory and runs prepare_walker on it.  Columns Ns..Wp-1 of the coefficient tile replay walker Ns-1.
#define SPP_THREADS 128

// ---- device-side exchange ordering of the sharded ensemble (replaces a separate barrier launch) -----------
// Every rank owns uint64 flags[world] in symmetric memory; flags[r] on rank k = the last half-step epoch whose
// position stores rank r has finished publishing into rank k's replica.
//   signal: the LAST CTA of the publishing kernel (ticket counter) stores the epoch into its slot on every
//           peer with a system-scope release (cumulative over the CTAs' stores, each followed by a
//           system fence before its ticket).
//   wait:   the first thing the next proposal kernel does -- thread r of every CTA spins on flags[r] with
//           acquire loads until it reaches the epoch, then the CTA proceeds to read the complementary half.
// A rank can therefore run ahead of a slow peer by at most one half-step, and the write-after-read hazard on the
// replica is excluded by the same chain (a peer's stores for half-step e+1 come after its wait for e, which
// comes after this rank's signal for e, which comes after this rank's reads for e).
// sync[0] = sticky status word (bit 0: a wait timed out -- the host raises), sync[1] = ticket counter.
__device__ __forceinline__ bool peer_wait_flag(const unsigned long long* mine, unsigned long long epoch,
                                               unsigned int* sync, long long timeout_cycles) {
  unsigned long long seen = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
    if (seen >= epoch) return true;
    if (sync && *reinterpret_cast<volatile unsigned int*>(sync) != 0u) return false;   // already failed: do not spin again
    if (clock64() - t0 > timeout_cycles) {
      if (sync) atomicOr(sync, 1u);
      else __trap();                       // no status word to report through: fail loudly
      return false;
    }
    __nanosleep(32);
  }
}

// all CTAs call this after their last peer store; the last one to arrive signals every peer
__device__ __forceinline__ void peer_signal_last_cta(const unsigned long long* __restrict__ flag_ptrs, int rank, int world,
                                                     unsigned long long epoch, unsigned int* sync, int debug = 0) {
  __shared__ unsigned int s_last;
  __syncthreads();                                  // this CTA's stores are all issued
  if (threadIdx.x == 0) {
    if (debug & 2) __threadfence(); else
    __threadfence_system();                         // ... and ordered before the ticket
    const unsigned int t = atomicAdd(sync + 1, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    if ((int)threadIdx.x < world) {
      __threadfence_system();
      unsigned long long* theirs = reinterpret_cast<unsigned long long*>(flag_ptrs[threadIdx.x]) + rank;
      asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(theirs), "l"(epoch) : "memory");
    }
    if (threadIdx.x == 0) sync[1] = 0u;             // ready for the next publishing launch (stream ordered)
  }
}

struct csl_peer_sync {
  const unsigned long long* flag_ptrs;   // device array [world] of every rank's flag array, or nullptr (one GPU)
  unsigned int* sync;                    // [2]: status, ticket
  unsigned long long epoch;              // wait: the epoch every peer must have reached; signal: the epoch announced
  long long timeout_cycles;
  int rank, world;
  int debug;                             // CSL_XCHG_DEBUG (timing experiments only; results are WRONG when set):
                                         // 1 no remote stores, 2 device-scope instead of system-scope CTA fence, 4 no wait
};

__global__ void __launch_bounds__(SPP_THREADS)
k_stretch_propose_prepare(csl_program_t p, int Ns, int Wp, int Nc, const double* __restrict__ s,
                          const double* comp, double a, uint64_t seed, uint32_t step, uint32_t half,
                          uint32_t walker0, double* __restrict__ q, double* __restrict__ zz,
                          double* __restrict__ coef, int ldc, double* __restrict__ prior, csl_peer_sync ps,
                          int stage_tables) {
  extern __shared__ __align__(16) double sq[];
  const int nd = p.ndim, lds = nd + 1;
  // everything that does not depend on the other ranks first: the bytecode / prior tables into shared memory and
  // the walker's Philox draws, all in flight while thread r polls peer r's flag
  const csl_small_tables tabs = stage_tables ? stage_small_tables(p, sq + SPP_THREADS * lds, threadIdx.x, SPP_THREADS)
                                             : small_tables_global(p);
  const int w = blockIdx.x * SPP_THREADS + threadIdx.x;
  const int k = min(w, Ns - 1);
  double uz, ua;
  int64_t j;
  philox_stretch(seed, walker0 + (uint32_t)k, half, step, (uint32_t)Nc, uz, j, ua);
  // zz = ((a - 1.) * u + 1) ** 2. / a
  const double tq = __dadd_rn(__dmul_rn(a - 1.0, uz), 1.0);
  const double z = __ddiv_rn(__dmul_rn(tq, tq), a);
  if (ps.flag_ptrs && !(ps.debug & 4)) { // sharded: the peers' stores of the previous half-step must have landed
    if ((int)threadIdx.x < ps.world)
      peer_wait_flag(reinterpret_cast<const unsigned long long*>(ps.flag_ptrs[ps.rank]) + threadIdx.x, ps.epoch,
                     ps.sync, ps.timeout_cycles);
  }
  __syncthreads();
  if (w >= Wp) return;
  const long long sb = (long long)k * nd, cb = (long long)j * nd;
  double* qr = sq + threadIdx.x * lds;
  // four components at a time: the partner-row loads of a group are independent and leave together
  for (int i0 = 0; i0 < nd; i0 += 4) {
    double c[4], sv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < nd) {
        c[u] = __ldcg(comp + cb + i0 + u);             // L2 is the point of coherence for the peers' stores
        sv[u] = s[sb + i0 + u];
      }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (i0 + u < nd) {
        const double v = __dsub_rn(c[u], __dmul_rn(z, __dsub_rn(c[u], sv[u])));     // q = c - zz * (c - s)
        qr[i0 + u] = v;
        if (w < Ns) q[sb + i0 + u] = v;
      }
  }
  if (w < Ns) zz[w] = z;
  prepare_walker(p, tabs, qr, w, Ns, coef, ldc, prior);
}

// Accept + publish.  With walkers sharded over several GPUs every rank keeps a replica of ALL walkers'
// positions (`gath`, symmetric memory).  The accept kernel itself stores each local walker's final
// position into every rank's replica over NVLink -- peer pointers (st.global on mapped peer memory) or
// one multimem.st through the NVSwitch multicast address -- so the "all-gather of the complementary
// half" is fused into the accept step; a device-side barrier follows on the stream.
__global__ void __launch_bounds__(SAMP_THREADS)
k_stretch_accept(int Ns, int nd, double* __restrict__ s, double* __restrict__ lnl_s, const double* __restrict__ q,
                 const double* __restrict__ lnl_q, const double* __restrict__ zz, const double* __restrict__ u_acc,
                 uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0, uint32_t Nc,
                 uint8_t* __restrict__ accepted, const unsigned long long* __restrict__ peers, int npeers,
                 unsigned long long mc_ptr, long long dst_row0, unsigned long long* __restrict__ nacc) {
  const int k = warp_global(), lane = threadIdx.x & 31;
  if (k >= Ns) {                               // whole warps only; keep them for the block-wide count
    if (nacc) __syncthreads_count(0);
    return;
  }
  double ua;
  if (u_acc) ua = u_acc[k];
  else { double uz; int64_t j; philox_stretch(seed, walker0 + (uint32_t)k, half, step, Nc, uz, j, ua); }
  const double lq = lnl_q[k];
  // lnpdiff = (dim - 1.) * log(zz) + newlnprob - lnprob0 ; accept = lnpdiff > log(u)
  const double newlnp = -lq, lnp0 = -lnl_s[k];
  const double lnpdiff = __dsub_rn(__dadd_rn(__dmul_rn((double)nd - 1.0, log(zz[k])), newlnp), lnp0);
  const bool acc = lnpdiff > log(ua);
  const long long b = (long long)k * nd;
  double v0 = 0.0, v1 = 0.0;
  if (lane < nd) v0 = acc ? q[b + lane] : s[b + lane];
  if (lane + 32 < nd) v1 = acc ? q[b + lane + 32] : s[b + lane + 32];
  if (acc) {
    if (lane < nd) s[b + lane] = v0;
    if (lane + 32 < nd) s[b + lane + 32] = v1;
    if (lane == 0) lnl_s[k] = lq;
  }
  if (lane == 0 && accepted) accepted[k] = acc ? 1 : 0;
  if (nacc) {                                   // one atomic per CTA
    const int c = __syncthreads_count(lane == 0 && acc);
    if (threadIdx.x == 0 && c) atomicAdd(nacc, (unsigned long long)c);
  }
  const long long drow = (dst_row0 + k) * nd;
  if (mc_ptr) {
    double* dst = reinterpret_cast<double*>(mc_ptr) + drow;
    if (lane < nd) asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(dst + lane), "d"(v0) : "memory");
    if (lane + 32 < nd) asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(dst + lane + 32), "d"(v1) : "memory");
  } else {
    for (int r = 0; r < npeers; ++r) {
      double* dst = reinterpret_cast<double*>(peers[r]) + drow;
      if (lane < nd) dst[lane] = v0;
      if (lane + 32 < nd) dst[lane + 32] = v1;
    }
  }
}

// Device-side barrier across the ranks of one NVLink domain, launched on the compute stream right after
// the publishing accept kernel of the STAGED (per-stage) path; the native step loop folds the same signal /
// wait into its accept and proposal kernels instead.  Thread r signals peer r with a system-scope release store
// (cumulative: the previous kernel's peer stores are ordered before it) and spins on its own slot r with
// acquire loads.  The host never blocks.  A wait that exceeds the timeout sets the sticky status word (the host
// checks it and raises) -- or traps when no status word was given; it never passes silently.
// signal = 0: wait only (end of csl_stretch_run: the caller may read the replica afterwards).
__global__ void k_peer_barrier(const unsigned long long* __restrict__ flag_ptrs, int rank, int world,
                               unsigned long long epoch, int signal, unsigned int* sync, long long timeout_cycles) {
  const int r = threadIdx.x;
  if (r >= world) return;
  if (signal) {
    __threadfence_system();
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(flag_ptrs[r]) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(theirs), "l"(epoch) : "memory");
  }
  peer_wait_flag(reinterpret_cast<const unsigned long long*>(flag_ptrs[rank]) + r, epoch, sync, timeout_cycles);
  __threadfence_system();
}

// Combine + accept (+ publish) of the native step loop: k_sum_partials, k_combine and k_stretch_accept in one
// launch.  Phase 1, one THREAD per walker: lnl(q) from the chi^2 partial sums (same order of additions as
// the staged kernels), the accept test, lnl update.  Phase 2, the CTA's 256 rows as one contiguous range:
// accepted rows are copied q -> s, and every row's final position goes to the replicas of all ranks.
struct csl_fused_rows {           // row bookkeeping folded into the accept kernel (pointers at this half's first walker)
  double* weight; double* rows; int32_t* nrows; int cap, last_step;
};
struct csl_fused_combine {
  csl_program_t p;
  const double* coef; int ldc;
  const double* prior; const double* chi2; const double* part; int nblk, ldp;
  double* lnl_q;
};
#define SAB_THREADS 256          // largest CTA; launches of a few thousand walkers use smaller ones (more CTAs)
__global__ void __launch_bounds__(SAB_THREADS)
k_stretch_combine_accept(int Ns, int nd, double* __restrict__ s, double* __restrict__ lnl_s,
                         const double* __restrict__ q, const double* __restrict__ zz, uint64_t seed, uint32_t step,
                         uint32_t half, uint32_t walker0, uint32_t Nc, uint8_t* __restrict__ accepted,
                         const unsigned long long* __restrict__ peers, int npeers, unsigned long long mc_ptr,
                         long long dst_row0, unsigned long long* __restrict__ nacc, csl_fused_combine fc,
                         csl_peer_sync ps, csl_fused_rows fr) {
  __shared__ unsigned char sacc[SAB_THREADS];
  const int nthr = blockDim.x;
  const int k0 = blockIdx.x * nthr, k = k0 + threadIdx.x;
  bool acc = false;
  if (k < Ns) {
    double uz, ua;
    int64_t j;
    philox_stretch(seed, walker0 + (uint32_t)k, half, step, Nc, uz, j, ua);
    double chi2 = 0.0;
    if (fc.part) for (int i = 0; i < fc.nblk; ++i) chi2 += fc.part[(long long)i * fc.ldp + k];
    else if (fc.chi2) chi2 = fc.chi2[k];
    const double lq = combine_walker(fc.p, k, fc.coef, fc.ldc, fc.prior[k], fc.part != nullptr || fc.chi2 != nullptr, chi2);
    fc.lnl_q[k] = lq;
    // lnpdiff = (dim - 1.) * log(zz) + newlnprob - lnprob0 ; accept = lnpdiff > log(u)
    const double newlnp = -lq, lnp0 = -lnl_s[k];
    const double lnpdiff = __dsub_rn(__dadd_rn(__dmul_rn((double)nd - 1.0, log(zz[k])), newlnp), lnp0);
    acc = lnpdiff > log(ua);
    if (acc) lnl_s[k] = lq;
    if (accepted) accepted[k] = acc ? 1 : 0;
    if (fr.rows) {
      // the wrapper's row bookkeeping (samplers/emcee.py:70-88) for this walker, whose position is final for this
      // step once its half is updated: unmoved -> weight + 1; moved (or last step) -> emit (lnl_next, weight, x_old).
      // k_emcee_rows compares the walker's previous position with the new one; here the previous position is s
      // itself (not yet overwritten: phase 2 below comes after a CTA barrier) and the new one is q if accepted.
      // (all components are loaded, four at a time: a short-circuit `&&` chained nd dependent L2 round trips)
      bool same = true;
      if (acc)
        for (int d0 = 0; d0 < nd; d0 += 4) {
          double a[4], b[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (d0 + u < nd) { a[u] = q[(long long)k * nd + d0 + u]; b[u] = s[(long long)k * nd + d0 + u]; }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (d0 + u < nd) same &= (a[u] == b[u]);
        }
      const double wt = fr.weight[k];
      if (same && !fr.last_step) {
        fr.weight[k] = wt + 1.0;
      } else {
        const int slot = fr.nrows[k];
        if (slot < fr.cap) {
          double* r = fr.rows + ((long long)k * fr.cap + slot) * (2 + nd);
          r[0] = acc ? lq : -lnp0;
          r[1] = wt;
          for (int d0 = 0; d0 < nd; d0 += 4) {
            double a[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (d0 + u < nd) a[u] = s[(long long)k * nd + d0 + u];
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (d0 + u < nd) r[2 + d0 + u] = a[u];
          }
        }
        fr.nrows[k] = slot + 1;                 // counts past cap signal overflow to the host
        fr.weight[k] = 1.0;
      }
    }
  }
  sacc[threadIdx.x] = acc ? 1 : 0;
  const int c = __syncthreads_count(acc);
  if (nacc && threadIdx.x == 0 && c) atomicAdd(nacc, (unsigned long long)c);
  const int nw = min(nthr, Ns - k0);
  const long long base = (long long)k0 * nd, dbase = (dst_row0 + k0) * nd;
  for (int idx = threadIdx.x; idx < nw * nd; idx += nthr) {
    const bool a = sacc[idx / nd] != 0;
    const double v = a ? q[base + idx] : s[base + idx];
    if (a) s[base + idx] = v;
    if (mc_ptr) {
      double* dst = reinterpret_cast<double*>(mc_ptr) + dbase + idx;
      asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(dst), "d"(v) : "memory");
    } else {
      for (int r = 0; r < npeers; ++r)
        if (!(ps.debug & 1) || r == ps.rank) reinterpret_cast<double*>(peers[r])[dbase + idx] = v;
    }
  }
  // sharded: announce this half-step to every peer once ALL CTAs of this launch have stored
  if (ps.flag_ptrs) peer_signal_last_cta(ps.flag_ptrs, ps.rank, ps.world, ps.epoch, ps.sync, ps.debug);
}

// Chain.add_gauss_prior on the device (chains.py:206-233): every stored row of every chain gets
//   dlnl = (x_P - mean)^T Cinv (x_P - mean) / 2 ,  weight *= exp(-dlnl) ,  lnl += dlnl
// for the np parameters listed in pidx (Cinv [np, np] row-major; a single parameter: Cinv = 1 / std^2).
// One thread per row slot; rows are (lnl, weight, x...) records of 2 + nd doubles.
__global__ void __launch_bounds__(256)
k_rows_gauss_prior(int W, int nd, double* __restrict__ rows, const int32_t* __restrict__ nrows, int cap, int np,
                   const int32_t* __restrict__ pidx, const double* __restrict__ mean, const double* __restrict__ cinv) {
  const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= (long long)W * cap) return;
  const int w = (int)(slot / cap), r = (int)(slot - (long long)w * cap);
  if (r >= min(nrows[w], cap)) return;
  double* row = rows + slot * (2 + nd);
  double q = 0.0;
  for (int i = 0; i < np; ++i) {
    const double di = row[2 + pidx[i]] - mean[i];
    double t = 0.0;
    for (int k = 0; k < np; ++k) t = fma(cinv[i * np + k], row[2 + pidx[k]] - mean[k], t);
    q = fma(di, t, q);
  }
  const double dlnl = 0.5 * q;
  row[1] *= exp(-dlnl);
  row[0] += dlnl;
}

// Chain.thin on the device (chains.py:235-243): one warp per chain walks its rows in order; with
// S_k = ceil(cumulative weight through row k / delta), row k is kept iff S_k > S_{k-1} and gets weight S_k - S_{k-1}.
// Kept rows are compacted to the front of the chain's buffer in place (the write index never passes the read index).
__global__ void __launch_bounds__(SAMP_THREADS)
k_rows_thin(int W, int nd, double* __restrict__ rows, int32_t* __restrict__ nrows, int cap, double delta) {
  const int w = warp_global(), lane = threadIdx.x & 31;
  if (w >= W) return;
  const int rl = 2 + nd, nr = min(nrows[w], cap);
  double* base = rows + (long long)w * cap * rl;
  double cum = 0.0, prev = 0.0;
  int out = 0;
  for (int r = 0; r < nr; ++r) {
    const double* src = base + (long long)r * rl;
    cum += src[1];
    const double slot = ceil(cum / delta);
    if (slot > prev) {                           // warp-uniform
      double* dst = base + (long long)out * rl;
      double v0 = 0.0, v1 = 0.0, v2 = 0.0;
      if (lane < rl) v0 = src[lane];
      if (lane + 32 < rl) v1 = src[lane + 32];
      if (lane + 64 < rl) v2 = src[lane + 64];
      __syncwarp();
      if (lane < rl) dst[lane] = (lane == 1) ? slot - prev : v0;
      if (lane + 32 < rl) dst[lane + 32] = v1;
      if (lane + 64 < rl) dst[lane + 64] = v2;
      __syncwarp();
      ++out;
      prev = slot;
    }
  }
  if (lane == 0) nrows[w] = out;
}

// wrapper bookkeeping, samplers/emcee.py:70-88
__global__ void __launch_bounds__(SAMP_THREADS)
k_emcee_rows(int W, int nd, double* __restrict__ pos_old, const double* __restrict__ pos_new,
             const double* __restrict__ lnl_new, double* __restrict__ weight, int last_step,
             double* __restrict__ rows, int32_t* __restrict__ nrows, int cap) {
  const int w = warp_global(), lane = threadIdx.x & 31;
  if (w >= W) return;
  const long long b = (long long)w * nd;
  double o0 = (lane < nd) ? pos_old[b + lane] : 0.0, o1 = (lane + 32 < nd) ? pos_old[b + lane + 32] : 0.0;
  double n0 = (lane < nd) ? pos_new[b + lane] : 0.0, n1 = (lane + 32 < nd) ? pos_new[b + lane + 32] : 0.0;
  const bool same = __all_sync(0xffffffffu, o0 == n0 && o1 == n1);
  double wt = weight[w];
  if (same && !last_step) {
    if (lane == 0) weight[w] = wt + 1.0;
  } else {
    emit_row(rows, nrows, cap, w, nd, lane, lnl_new[w], wt, o0, o1);
    if (lane == 0) weight[w] = 1.0;
    // this step's positions are the next step's "old" ones
    if (lane < nd) pos_old[b + lane] = n0;
    if (lane + 32 < nd) pos_old[b + lane + 32] = n1;
  }
}

// per-chain weighted count / mean / variance (denominator sum w - 1) over rows [burn, nrows)
__global__ void __launch_bounds__(SAMP_THREADS)
k_chain_moments(int W, int nd, const double* __restrict__ rows, const int32_t* __restrict__ nrows, int cap,
                int burn, double* __restrict__ stats) {
  const int w = warp_global(), lane = threadIdx.x & 31;
  if (w >= W) return;
  const int rl = 2 + nd, nr = min(nrows[w], cap);
  const double* base = rows + (long long)w * cap * rl;
  double sw = 0.0, m0 = 0.0, m1 = 0.0;
  for (int r = burn; r < nr; ++r) {
    const double wt = base[(long long)r * rl + 1];
    sw += wt;
    if (lane < nd) m0 = fma(wt, base[(long long)r * rl + 2 + lane], m0);
    if (lane + 32 < nd) m1 = fma(wt, base[(long long)r * rl + 2 + lane + 32], m1);
  }
  m0 /= sw; m1 /= sw;
  double v0 = 0.0, v1 = 0.0;
  for (int r = burn; r < nr; ++r) {
    const double wt = base[(long long)r * rl + 1];
    if (lane < nd) { double d = base[(long long)r * rl + 2 + lane] - m0; v0 = fma(wt * d, d, v0); }
    if (lane + 32 < nd) { double d = base[(long long)r * rl + 2 + lane + 32] - m1; v1 = fma(wt * d, d, v1); }
  }
  double* o = stats + (long long)w * (1 + 2 * nd);
  if (lane == 0) o[0] = sw;
  if (lane < nd) { o[1 + lane] = m0; o[1 + nd + lane] = v0 / (sw - 1.0); }
  if (lane + 32 < nd) { o[1 + lane + 32] = m1; o[1 + nd + lane + 32] = v1 / (sw - 1.0); }
}

// test hooks: the device streams, to be compared with oracle/philox.py
__global__ void k_philox_mh(int W, int nd, uint64_t seed, uint32_t step, uint32_t chain0, double* z, double* u) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  for (int pr = 0; 2 * pr < nd; ++pr) {
    double z0, z1;
    philox_normal_pair(seed, chain0 + (uint32_t)w, (uint32_t)pr, step, z0, z1);
    z[(long long)w * nd + 2 * pr] = z0;
    if (2 * pr + 1 < nd) z[(long long)w * nd + 2 * pr + 1] = z1;
  }
  u[w] = philox_mh_uniform(seed, chain0 + (uint32_t)w, step);
}
__global__ void k_philox_stretch(int Ns, int Nc, uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0,
                                 double* u_z, int64_t* j, double* u_acc) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= Ns) return;
  philox_stretch(seed, walker0 + (uint32_t)k, half, step, (uint32_t)Nc, u_z[k], j[k], u_acc[k]);
}

// =============================================================================================
static inline int warp_grid(int nwarps) { return (nwarps + SAMP_WARPS - 1) / SAMP_WARPS; }

extern "C" int csl_mh_propose(int W, int ndim, const double* cur_x, double* test_x, int mode, const double* inj,
                              const double* F, uint64_t seed, uint32_t step, uint32_t chain0, void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && cur_x && test_x, "csl_mh_propose: bad argument");
  CSL_REQUIRE(mode >= 0 && mode <= 2, "csl_mh_propose: mode must be 0, 1 or 2");
  CSL_REQUIRE(mode == 2 || inj, "csl_mh_propose: injected stream missing");
  CSL_REQUIRE(mode == 0 || F, "csl_mh_propose: proposal factor missing");
  static_assert((size_t)CSL_MAX_DIM * CSL_MAX_DIM * sizeof(double) <= 48 * 1024, "proposal factor must fit the default dynamic shared memory");
  size_t smem = mode ? (size_t)ndim * ndim * sizeof(double) : 0;
  k_mh_propose<<<warp_grid(W), SAMP_THREADS, smem, (cudaStream_t)stream>>>(W, ndim, cur_x, test_x, mode, inj, F,
                                                                           seed, step, chain0);
  return csl_check_launch("k_mh_propose");
}

extern "C" int csl_mh_accept(int W, int ndim, double* cur_x, double* cur_lnl, double* cur_weight,
                             const double* test_x, const double* test_lnl, const double* u, uint64_t seed,
                             uint32_t step, uint32_t chain0, double temp, double max_weight, double* rows,
                             int32_t* nrows, int cap, uint8_t* accepted, double* margin, void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM, "csl_mh_accept: bad W / ndim");
  CSL_REQUIRE(cur_x && cur_lnl && cur_weight && test_x && test_lnl && rows && nrows && cap > 0,
              "csl_mh_accept: null argument");
  k_mh_accept<<<warp_grid(W), SAMP_THREADS, 0, (cudaStream_t)stream>>>(W, ndim, cur_x, cur_lnl, cur_weight, test_x,
                                                                       test_lnl, u, seed, step, chain0, temp,
                                                                       max_weight, rows, nrows, cap, accepted, margin);
  return csl_check_launch("k_mh_accept");
}

extern "C" int csl_mh_gauss_run(const csl_program_t* p, int W, int nsteps, const double* Lfac, const double* mean,
                                const int32_t* perm, double* cur_x, double* cur_lnl, double* cur_weight,
                                const double* z, int z_is_delta, const double* u, const double* F, uint64_t seed,
                                uint32_t step0, uint32_t chain0, double temp, double max_weight, double* rows,
                                int32_t* nrows, int cap, void* stream) {
  CSL_REQUIRE(p && W > 0 && nsteps > 0 && Lfac && mean && perm && cur_x && cur_lnl && cur_weight && rows && nrows,
              "csl_mh_gauss_run: null argument");
  CSL_REQUIRE(p->ndim >= 1 && p->ndim <= 32 && p->n >= 1 && p->n <= 32, "csl_mh_gauss_run: ndim and n must be <= 32");
  CSL_REQUIRE((z && z_is_delta) || F, "csl_mh_gauss_run: proposal factor missing");
  static_assert((size_t)(32 * 32 + 32 * 32 + 32) * sizeof(double) <= 48 * 1024, "factors of the 32-parameter limit above fit the default dynamic shared memory");
  size_t smem = (size_t)(p->ndim * p->ndim + p->n * p->n + p->n) * sizeof(double);
  k_mh_gauss_run<<<warp_grid(W), SAMP_THREADS, smem, (cudaStream_t)stream>>>(
      *p, W, nsteps, Lfac, mean, perm, cur_x, cur_lnl, cur_weight, z, z_is_delta, u, F, seed, step0, chain0, temp,
      max_weight, rows, nrows, cap);
  return csl_check_launch("k_mh_gauss_run");
}

static int moments_impl(int W, int ndim, const double* rows, const int32_t* nrows, int cap, const double* mean,
                        double* partial, double* out, int burn, void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && rows && nrows && partial && out, "csl_mh_moments: bad argument");
  static_assert((size_t)MOM_TILE * (2 + CSL_MAX_DIM) * sizeof(double) <= 48 * 1024, "row tile must fit the default dynamic shared memory");
  const int m = 1 + ndim + ndim * ndim;
  cudaStream_t s = (cudaStream_t)stream;
  k_mh_moments<<<MOM_BLOCKS, SAMP_THREADS, (size_t)MOM_TILE * (2 + ndim) * sizeof(double), s>>>(W, ndim, rows, nrows, cap, mean, partial, burn);
  int lo = mean ? 1 + ndim : 0, hi = mean ? m : 1 + ndim;
  k_reduce_partials<<<(hi - lo + 127) / 128, 128, 0, s>>>(m, lo, hi, partial, out);
  return csl_check_launch("k_mh_moments");
}

extern "C" int csl_mh_moments(int W, int ndim, const double* rows, const int32_t* nrows, int cap,
                              const double* mean, double* partial, double* out, void* stream) {
  return moments_impl(W, ndim, rows, nrows, cap, mean, partial, out, -1, stream);
}

extern "C" int csl_rows_moments(int W, int ndim, const double* rows, const int32_t* nrows, int cap, int burn_rows,
                                const double* mean, double* partial, double* out, void* stream) {
  CSL_REQUIRE(burn_rows >= 0, "csl_rows_moments: burn_rows must be >= 0");
  return moments_impl(W, ndim, rows, nrows, cap, mean, partial, out, burn_rows, stream);
}

extern "C" int csl_rows_gauss_prior(int W, int ndim, double* rows, const int32_t* nrows, int cap, int np,
                                    const int32_t* pidx, const double* mean, const double* cinv, void* stream) {
  CSL_REQUIRE(W > 0 && cap > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && rows && nrows, "csl_rows_gauss_prior: bad argument");
  CSL_REQUIRE(np >= 1 && np <= ndim && pidx && mean && cinv, "csl_rows_gauss_prior: parameter list");
  const long long slots = (long long)W * cap;
  k_rows_gauss_prior<<<(unsigned)((slots + 255) / 256), 256, 0, (cudaStream_t)stream>>>(W, ndim, rows, nrows, cap, np,
                                                                                       pidx, mean, cinv);
  return csl_check_launch("k_rows_gauss_prior");
}

extern "C" int csl_rows_thin(int W, int ndim, double* rows, int32_t* nrows, int cap, double delta, void* stream) {
  CSL_REQUIRE(W > 0 && cap > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && rows && nrows && delta > 0, "csl_rows_thin: bad argument");
  k_rows_thin<<<warp_grid(W), SAMP_THREADS, 0, (cudaStream_t)stream>>>(W, ndim, rows, nrows, cap, delta);
  return csl_check_launch("k_rows_thin");
}

extern "C" int csl_stretch_propose(int Ns, int Nc, int ndim, const double* s, const double* comp, double a,
                                   const double* u_z, const int64_t* j, uint64_t seed, uint32_t step, uint32_t half,
                                   uint32_t walker0, double* q, double* zz, void* stream) {
  CSL_REQUIRE(Ns > 0 && Nc > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && s && comp && q && zz,
              "csl_stretch_propose: bad argument");
  CSL_REQUIRE((u_z == nullptr) == (j == nullptr), "csl_stretch_propose: u_z and j must be injected together");
  k_stretch_propose<<<warp_grid(Ns), SAMP_THREADS, 0, (cudaStream_t)stream>>>(Ns, Nc, ndim, s, comp, a, u_z, j, seed,
                                                                             step, half, walker0, q, zz);
  return csl_check_launch("k_stretch_propose");
}

static int stretch_accept_impl(int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q, const double* lnl_q,
                               const double* zz, const double* u_acc, uint64_t seed, uint32_t step, uint32_t half,
                               uint32_t walker0, uint8_t* accepted, const unsigned long long* peers, int npeers,
                               unsigned long long mc_ptr, long long dst_row0, unsigned long long* nacc, void* stream) {
  CSL_REQUIRE(Ns > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && s && lnl_s && q && lnl_q && zz,
              "csl_stretch_accept: bad argument");
  CSL_REQUIRE(npeers == 0 || peers || mc_ptr, "csl_stretch_accept: peer pointers missing");
  k_stretch_accept<<<warp_grid(Ns), SAMP_THREADS, 0, (cudaStream_t)stream>>>(
      Ns, ndim, s, lnl_s, q, lnl_q, zz, u_acc, seed, step, half, walker0, (uint32_t)Nc, accepted, peers, npeers, mc_ptr,
      dst_row0, nacc);
  return csl_check_launch("k_stretch_accept");
}

// ---- fused stages of the native step loop (api.cu: csl_stretch_run); not part of the C ABI -----------------
long long csl_barrier_timeout_cycles() {
  // CSL_BARRIER_TIMEOUT_S (default 120 s): how long a rank waits for a peer's half-step before it gives up.
  // Read once; clock64 ticks at the SM clock (<= 2 GHz on B200).
  static long long cached = 0;
  if (!cached) {
    double sec = 120.0;
    if (const char* e = getenv("CSL_BARRIER_TIMEOUT_S")) { const double v = atof(e); if (v > 0) sec = v; }
    cached = (long long)(sec * 2.0e9);
  }
  return cached;
}

static csl_peer_sync make_peer_sync(const csl_peer_args* a) {
  csl_peer_sync ps;
  ps.flag_ptrs = a ? reinterpret_cast<const unsigned long long*>(a->flag_ptrs) : nullptr;
  ps.sync = a ? a->sync : nullptr;
  ps.epoch = a ? (unsigned long long)a->epoch : 0ull;
  ps.timeout_cycles = csl_barrier_timeout_cycles();
  ps.rank = a ? a->rank : 0;
  ps.world = a ? a->world : 0;
  static const int dbg = csl_env_int("CSL_XCHG_DEBUG", 0);
  ps.debug = dbg;
  return ps;
}

int csl_fused_propose_prepare(const csl_program_t* p, int Ns, int Wp, int Nc, const double* s, const double* comp,
                              double a, uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0, double* q,
                              double* zz, double* coef, int ldc, double* prior, const csl_peer_args* wait,
                              void* stream) {
  size_t smem = (size_t)SPP_THREADS * (p->ndim + 1) * sizeof(double);
  const int stage_tables = csl_small_stage_bytes(p) > 0 ? 1 : 0;
  smem += csl_small_stage_bytes(p);
  if (smem > 48 * 1024) {                      // above the default from ndim = 48 on; per device, cheap: set every time
    cudaError_t e = cudaFuncSetAttribute(k_stretch_propose_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  k_stretch_propose_prepare<<<(Wp + SPP_THREADS - 1) / SPP_THREADS, SPP_THREADS, smem, (cudaStream_t)stream>>>(
      *p, Ns, Wp, Nc, s, comp, a, seed, step, half, walker0, q, zz, coef, ldc, prior, make_peer_sync(wait), stage_tables);
  return csl_check_launch("k_stretch_propose_prepare");
}

int csl_fused_combine_accept(const csl_program_t* p, int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q,
                             double* lnl_q, const double* zz, uint64_t seed, uint32_t step, uint32_t half,
                             uint32_t walker0, uint8_t* accepted, uint64_t* naccepted, const uint64_t* peer_ptrs,
                             int npeers, uint64_t multicast_ptr, int64_t dst_row0, const double* coef, int ldc,
                             const double* prior, const double* chi2, const double* part, int nblk, int ldp,
                             const csl_peer_args* signal, double* weight, double* rows, int32_t* nrows, int cap,
                             int last_step, void* stream) {
  csl_fused_rows fr;
  fr.weight = weight; fr.rows = rows; fr.nrows = nrows; fr.cap = cap; fr.last_step = last_step;
  csl_fused_combine fc;
  fc.p = *p; fc.coef = coef; fc.ldc = ldc; fc.prior = prior; fc.chi2 = chi2; fc.part = part; fc.nblk = nblk;
  fc.ldp = ldp; fc.lnl_q = lnl_q;
  // CTA size: 256 walkers per CTA in bulk; launches of a few thousand walkers take 64 so that the serial
  // per-walker phase and the peer stores spread over more SMs (CSL_SAB_THREADS overrides, for sweeps)
  const int forced = csl_option_value("sab_threads");
  int nthr = forced > 0 ? forced : (Ns <= 16384 ? 64 : SAB_THREADS);
  nthr = nthr < 32 ? 32 : (nthr > SAB_THREADS ? SAB_THREADS : nthr / 32 * 32);
  k_stretch_combine_accept<<<(Ns + nthr - 1) / nthr, nthr, 0, (cudaStream_t)stream>>>(
      Ns, ndim, s, lnl_s, q, zz, seed, step, half, walker0, (uint32_t)Nc, accepted,
      reinterpret_cast<const unsigned long long*>(peer_ptrs), npeers, (unsigned long long)multicast_ptr,
      (long long)dst_row0, reinterpret_cast<unsigned long long*>(naccepted), fc, make_peer_sync(signal), fr);
  return csl_check_launch("k_stretch_combine_accept");
}

extern "C" int csl_stretch_accept(int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q,
                                  const double* lnl_q, const double* zz, const double* u_acc, uint64_t seed,
                                  uint32_t step, uint32_t half, uint32_t walker0, uint8_t* accepted,
                                  uint64_t* naccepted, void* stream) {
  return stretch_accept_impl(Ns, Nc, ndim, s, lnl_s, q, lnl_q, zz, u_acc, seed, step, half, walker0, accepted, nullptr,
                             0, 0ull, 0, reinterpret_cast<unsigned long long*>(naccepted), stream);
}

extern "C" int csl_stretch_accept_publish(int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q,
                                          const double* lnl_q, const double* zz, const double* u_acc, uint64_t seed,
                                          uint32_t step, uint32_t half, uint32_t walker0, uint8_t* accepted,
                                          uint64_t* naccepted, const uint64_t* peer_ptrs, int npeers,
                                          uint64_t multicast_ptr, int64_t dst_row0, void* stream) {
  CSL_REQUIRE(npeers > 0 && (peer_ptrs || multicast_ptr), "csl_stretch_accept_publish: no peers given");
  return stretch_accept_impl(Ns, Nc, ndim, s, lnl_s, q, lnl_q, zz, u_acc, seed, step, half, walker0, accepted,
                             reinterpret_cast<const unsigned long long*>(peer_ptrs), npeers,
                             (unsigned long long)multicast_ptr, (long long)dst_row0,
                             reinterpret_cast<unsigned long long*>(naccepted), stream);
}

static int peer_barrier_impl(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, int signal,
                             uint32_t* sync, void* stream) {
  CSL_REQUIRE(flag_ptrs && world > 0 && world <= 32 && rank >= 0 && rank < world, "csl_peer_barrier: bad argument");
  k_peer_barrier<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(flag_ptrs), rank, world,
                                                      (unsigned long long)epoch, signal, sync,
                                                      csl_barrier_timeout_cycles());
  return csl_check_launch("k_peer_barrier");
}

extern "C" int csl_peer_barrier(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, uint32_t* sync,
                                void* stream) {
  return peer_barrier_impl(flag_ptrs, rank, world, epoch, 1, sync, stream);
}

extern "C" int csl_peer_wait(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, uint32_t* sync,
                             void* stream) {
  return peer_barrier_impl(flag_ptrs, rank, world, epoch, 0, sync, stream);
}

// a rank with no walker in a half still owes its peers the half-step's signal
__global__ void k_peer_signal(const unsigned long long* __restrict__ flag_ptrs, int rank, int world,
                              unsigned long long epoch) {
  const int r = threadIdx.x;
  if (r >= world) return;
  __threadfence_system();
  unsigned long long* theirs = reinterpret_cast<unsigned long long*>(flag_ptrs[r]) + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(theirs), "l"(epoch) : "memory");
}

int csl_peer_signal(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, void* stream) {
  CSL_REQUIRE(flag_ptrs && world > 0 && world <= 32 && rank >= 0 && rank < world, "csl_peer_signal: bad argument");
  k_peer_signal<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long*>(flag_ptrs), rank, world,
                                                     (unsigned long long)epoch);
  return csl_check_launch("k_peer_signal");
}

extern "C" int csl_emcee_rows(int W, int ndim, double* pos_old, const double* pos_new, const double* lnl_new,
                              double* weight, int last_step, double* rows, int32_t* nrows, int cap, void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && pos_old && pos_new && lnl_new && weight && rows && nrows,
              "csl_emcee_rows: bad argument");
  k_emcee_rows<<<warp_grid(W), SAMP_THREADS, 0, (cudaStream_t)stream>>>(W, ndim, pos_old, pos_new, lnl_new, weight,
                                                                       last_step, rows, nrows, cap);
  return csl_check_launch("k_emcee_rows");
}

extern "C" int csl_chain_moments(int W, int ndim, const double* rows, const int32_t* nrows, int cap, int burn_rows,
                                 double* stats, void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && ndim <= CSL_MAX_DIM && rows && nrows && stats, "csl_chain_moments: bad argument");
  k_chain_moments<<<warp_grid(W), SAMP_THREADS, 0, (cudaStream_t)stream>>>(W, ndim, rows, nrows, cap, burn_rows, stats);
  return csl_check_launch("k_chain_moments");
}

extern "C" int csl_philox_mh(int W, int ndim, uint64_t seed, uint32_t step, uint32_t chain0, double* z, double* u,
                             void* stream) {
  CSL_REQUIRE(W > 0 && ndim >= 1 && z && u, "csl_philox_mh: bad argument");
  k_philox_mh<<<(W + 127) / 128, 128, 0, (cudaStream_t)stream>>>(W, ndim, seed, step, chain0, z, u);
  return csl_check_launch("k_philox_mh");
}

extern "C" int csl_philox_stretch(int Ns, int Nc, uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0,
                                  double* u_z, int64_t* j, double* u_acc, void* stream) {
  CSL_REQUIRE(Ns > 0 && Nc > 0 && u_z && j && u_acc, "csl_philox_stretch: bad argument");
  k_philox_stretch<<<(Ns + 127) / 128, 128, 0, (cudaStream_t)stream>>>(Ns, Nc, seed, step, half, walker0, u_z, j,
                                                                      u_acc);
  return csl_check_launch("k_philox_stretch");
}
