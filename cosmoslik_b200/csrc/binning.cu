// Foreground build + window binning + residual, one kernel.
//   reference: spt_lowl.py:171-182 (cl = cmb + egfs; b_i = <cl[windowrange], w_i>), :132 (r = cal*d - b),
//              spt_lowl.py:206-207 / models/clust_poisson_egfs.py:14-17 (the foreground spectra),
//              mspec_lnl.py:50-73 (the same per cross-spectrum).
// C[bins, walkers] = Wnd[bins, ell] * Cl[ell, walkers] runs on the FP64 tensor pipe (DMMA.8x8x4).
// The Cl operand never exists in HBM: each CTA builds its [16 ell x 128 walker] chunk in shared
// memory from the walker's coefficients (registers) and the spectrum's ell-table, whose 16 rows for
// the chunk are staged in shared memory by cp.async together with the window tile.
//
// ell-table row (doubles): [T_0 .. T_{NT-1}, (logB_q, M_q, D_q, 0) for q < NP],  D_q[l] = logB_q[l+1] - logB_q[l]
//   cl[l] = sum_t coef_t T_t[l] + sum_q amp_q M_q[l] exp(idx_q logB_q[l])
// The kernel is instantiated per padded class <NT, NP> so the build loop is branch-free and fully
// unrolled; the host sorts the tiles by class (csl_program_t::cls).
#include "common.cuh"
#include <type_traits>

// exp(y) for small |y| by its Taylor polynomial: degree 6 for |y| <= 0.012 (truncation < 7e-18), degree 10
// for |y| <= 0.1 (< 3e-19).  Used to advance a power law E[l] = exp(idx logB[l]) to the next column:
// E[l+1] = E[l] exp(idx D[l]),  D[l] = logB[l+1] - logB[l] from the ell-table.
// 1/k! in constant memory: DFMA takes the coefficient as a constant-bank operand, no moves needed
__constant__ double c_inv_fact[11] = {1.0, 1.0, 0.5, 1.6666666666666666e-01, 4.1666666666666664e-02,
                                      8.3333333333333332e-03, 1.3888888888888889e-03, 1.9841269841269841e-04,
                                      2.4801587301587302e-05, 2.7557319223985888e-06, 2.7557319223985893e-07};
template <int DEG>
__device__ __forceinline__ double exp_taylor(double y) {
  double p = c_inv_fact[DEG];
#pragma unroll
  for (int k = DEG - 1; k >= 0; --k) p = fma(p, y, c_inv_fact[k]);
  return p;
}

#define BW_XLD (CSL_BN + 2)       // row stride of a parked partial tile (doubles)
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem(const double* local_smem, uint32_t rank) {
  uint32_t remote;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_smem)), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
  return v;
}


template <int NT, int NP>
__global__ void __launch_bounds__(CSL_THREADS, 2)
k_bin_residual(csl_program_t p, int tile0, int W, int nsplit, const double* __restrict__ coef, int ldc,
               double* __restrict__ R, int ldr) {
  // nsplit > 1: launched as thread-block clusters of nsplit CTAs along x; CTA r of a cluster takes the r-th part
  // of the tile's ell chunks and the partial tiles are summed in rank order through distributed shared memory
  // (see k_bin_residual_ws below for the full story; nsplit is a property of the PROGRAM, csl_program_t::bin_split)
  constexpr int LSTRIDE = NT + 4 * NP;
  extern __shared__ __align__(128) double smem[];
  double* As = smem;                       // [2][CSL_ASZ]
  double* Bs = smem + 2 * CSL_ASZ;         // [2][CSL_BSZ]
  double* Ls = Bs + 2 * CSL_BSZ;           // [2][CSL_KC * LSTRIDE]
  const int split = (int)(blockIdx.x % (unsigned)nsplit);
  const int32_t* tile = p.tile_tab + (long long)(tile0 + blockIdx.x / nsplit) * CSL_TILE_STRIDE;
  const int32_t* spec = p.spec_tab + (long long)tile[CSL_T_SPEC] * CSL_SPEC_STRIDE;
  const int c0 = blockIdx.y * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;

  int l_begin = tile[CSL_T_LBEGIN], l_end = tile[CSL_T_LEND];
  if (nsplit > 1) {      // this CTA's contiguous share of the chunks; the first (n % nsplit) shares are one longer
    const int nall = (l_end - l_begin) / CSL_KC, cb = nall / nsplit, ce = nall % nsplit;
    l_begin += (split * cb + min(split, ce)) * CSL_KC;
    l_end = l_begin + (cb + (split < ce ? 1 : 0)) * CSL_KC;
  }
  const long long win_off = ((long long)tile[CSL_T_WINOFF_HI] << 32) | (unsigned)tile[CSL_T_WINOFF_LO];
  const long long win_ld = tile[CSL_T_WINLD];
  const double* __restrict__ Wnd = p.windows + win_off;
  const double* __restrict__ ltab = p.templates + spec[CSL_S_LTAB];

  // builder role: this thread owns walker c0 + wl and 8 of the 16 ell rows of a chunk
  const int wl = tid & (CSL_BN - 1), grp = tid >> 7;
  double cf[NT > 0 ? NT : 1], pamp[NP > 0 ? NP : 1], pidx[NP > 0 ? NP : 1];
#pragma unroll
  for (int t = 0; t < NT; ++t) cf[t] = coef[(long long)spec[CSL_S_TERM_COEF + t] * ldc + c0 + wl];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    pamp[q] = coef[(long long)spec[CSL_S_PLAW_AMP + q] * ldc + c0 + wl];
    pidx[q] = coef[(long long)spec[CSL_S_PLAW_IDX + q] * ldc + c0 + wl];
  }

  // stage the 16 ell-table rows of a chunk (contiguous 16*LSTRIDE doubles) with 16-byte cp.async
  auto load_L = [&](double* L, int l0) {
    constexpr int NCH = CSL_KC * LSTRIDE / 2;
    for (int c = tid; c < NCH; c += CSL_THREADS) cp_async16(L + 2 * c, ltab + (long long)l0 * LSTRIDE + 2 * c);
  };
  // the host's bound on |idx_q| (spectrum record; negative = hard bound from a prior box): the Taylor degree of
  // a column is chosen from bound * |D|, the same for every walker, so a walker's value never depends on its
  // neighbours; a walker outside a soft bound takes exp at every column
  double bnd[NP > 0 ? NP : 1];
  bool inside[NP > 0 ? NP : 1];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const float bq = __int_as_float(spec[CSL_S_IDXBOUND + q]);
    bnd[q] = fabs((double)bq);
    inside[q] = bq < 0.f || fabs(pidx[q]) <= bnd[q];
  }
  // 8 consecutive columns per thread: one exact exp at the first, Taylor advances after it (a column with a
  // large log-step falls back to exp)
  // aberration (SPTSZ_lowl.py:78-79, 84-89): cl[l] *= 1 + G[l] (ln cl[l] - ln cl[l-1]); G = 0 at the first column and
  // on the padding.  The value of the column before a thread's first one is evaluated from the table in global memory.
  const bool aberr = spec[CSL_S_ABERR] != 0;
  const double* __restrict__ gtab = p.templates + spec[CSL_S_ABTAB];
  auto build_B = [&](double* B, const double* L, int lchunk) {
    const double* row = L + grp * (CSL_KC / 2) * LSTRIDE;     // warp-uniform address -> shared-memory broadcast
    double E[NP > 0 ? NP : 1];
#pragma unroll
    for (int q = 0; q < NP; ++q) E[q] = exp(pidx[q] * row[NT + 4 * q]);
    const int lcol = lchunk + grp * (CSL_KC / 2);
    double lvprev = 0.0;
    if (aberr && lcol > 0) {
      const double* rp = ltab + (long long)(lcol - 1) * LSTRIDE;
      double vp = 0.0;
#pragma unroll
      for (int t = 0; t < NT; ++t) vp = fma(cf[t], __ldg(rp + t), vp);
#pragma unroll
      for (int q = 0; q < NP; ++q) vp = fma(pamp[q] * __ldg(rp + NT + 4 * q + 1), exp(pidx[q] * __ldg(rp + NT + 4 * q)), vp);
      lvprev = log(vp);
    }
#pragma unroll 2
    for (int r = 0; r < CSL_KC / 2; ++r, row += LSTRIDE) {
      double v = 0.0;
#pragma unroll
      for (int t = 0; t < NT; ++t) v = fma(cf[t], row[t], v);
#pragma unroll
      for (int q = 0; q < NP; ++q) v = fma(pamp[q] * row[NT + 4 * q + 1], E[q], v);
      if (aberr) {
        const double gk = __ldg(gtab + lcol + r), lv = log(v);
        if (gk != 0.0) v = v * (1.0 + gk * (lv - lvprev));
        lvprev = lv;
      }
      B[(grp * (CSL_KC / 2) + r) * CSL_LDB + wl] = v;
      if (r + 1 < CSL_KC / 2) {
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          const double D = row[NT + 4 * q + 2];
          const double y = pidx[q] * D, ay = bnd[q] * fabs(D);      // ay: warp-uniform (broadcast row)
          if (ay <= 0.012 && inside[q]) E[q] *= exp_taylor<6>(y);
          else if (ay <= 0.1 && inside[q]) E[q] *= exp_taylor<10>(y);
          else E[q] = exp(pidx[q] * row[LSTRIDE + NT + 4 * q]);
        }
      }
    }
  };

  // per 8-row group non-zero column range -> skip all-zero DMMA blocks (banded windows)
  int glo[4], ghi[4];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi) {
    glo[mi] = tile[CSL_T_GLO + (wm0 >> 3) + mi];
    ghi[mi] = tile[CSL_T_GHI + (wm0 >> 3) + mi];
  }

  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

  const int nchunk = (l_end - l_begin) / CSL_KC;
  if (nchunk > 0) {
    // prologue: table rows of chunks 0 and 1, window tile of chunk 0; build B(0)
    load_L(Ls, l_begin);
    load_A_tile_async(As, Wnd + l_begin, win_ld, tid);
    cp_async_commit();
    if (nchunk > 1) load_L(Ls + CSL_KC * LSTRIDE, l_begin + CSL_KC);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    build_B(Bs, Ls, l_begin);
    cp_async_wait<0>();
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
      const int cur = c & 1;
      const int l0 = l_begin + c * CSL_KC;
      // in flight during this iteration: window tile of chunk c+1, table rows of chunk c+2
      // (Ls[cur] held chunk c's rows, consumed one iteration ago)
      if (c + 1 < nchunk) load_A_tile_async(As + (cur ^ 1) * CSL_ASZ, Wnd + l0 + CSL_KC, win_ld, tid);
      if (c + 2 < nchunk) load_L(Ls + cur * CSL_KC * LSTRIDE, l0 + 2 * CSL_KC);
      cp_async_commit();
      if (c + 1 < nchunk) build_B(Bs + (cur ^ 1) * CSL_BSZ, Ls + (cur ^ 1) * CSL_KC * LSTRIDE, l0 + CSL_KC);
      unsigned mmask = 0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
        if (l0 + CSL_KC > glo[mi] && l0 < ghi[mi]) mmask |= 1u << mi;
      if (mmask) warp_mma_chunk(acc, As + cur * CSL_ASZ, Bs + cur * CSL_BSZ, wm0, wn0, lane, mmask);
      cp_async_wait<0>();
      __syncthreads();
    }
  }

  // epilogue: r = cal * d - b   (spt_lowl.py:132)
  const int g = lane >> 2, t = lane & 3;
  const int row0 = tile[CSL_T_ROW0], nrows = tile[CSL_T_NROWS];
  const int calrow = spec[CSL_S_CAL];
  const bool cal_on_model = spec[CSL_S_CALMODE] != 0;      // r = d - cal*b  (SPTSZ_lowl.py:68)
  if (nsplit > 1) {
    // park the partial tile over the (now idle) stages, meet the cluster's other parts in DSMEM: CTA r sums rows
    // [64 r / nsplit, 64 (r + 1) / nsplit) of all parts in rank order and writes their residuals
    double* X = smem;                                       // [64][BW_XLD]
    __syncthreads();                                        // every warp has left the stages
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        double2 v;
        v.x = acc[mi][ni][0];
        v.y = acc[mi][ni][1];
        *reinterpret_cast<double2*>(X + (wm0 + mi * 8 + g) * BW_XLD + wn0 + ni * 8 + 2 * t) = v;
      }
    cluster_sync_all();
    const int rows_per = CSL_BM / nsplit, r_lo = split * rows_per;
    for (int e = tid; e < rows_per * CSL_BN; e += CSL_THREADS) {
      const int rl = r_lo + e / CSL_BN, cl = e % CSL_BN;
      const double* src = X + rl * BW_XLD + cl;
      double b = ld_dsmem(src, 0);
      for (int q = 1; q < nsplit; ++q) b += ld_dsmem(src, (uint32_t)q);      // rank order: fixed
      if (rl < nrows) {
        const int col = c0 + cl;
        const double cal = calrow >= 0 ? coef[(long long)calrow * ldc + col] : 1.0;
        const double d = p.data[row0 + rl];
        R[(long long)(row0 + rl) * ldr + col] = cal_on_model ? __dsub_rn(d, __dmul_rn(b, cal)) : __dsub_rn(__dmul_rn(cal, d), b);
      }
    }
    cluster_sync_all();                                     // nobody leaves while a peer still reads its tile
    return;
  }
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) {
    const int col = c0 + wn0 + ni * 8 + 2 * t;
    double cal0 = 1.0, cal1 = 1.0;
    if (calrow >= 0) {
      cal0 = coef[(long long)calrow * ldc + col];
      cal1 = coef[(long long)calrow * ldc + col + 1];
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int rl = wm0 + mi * 8 + g;
      if (rl < nrows) {
        const double d = p.data[row0 + rl];
        double2 out;
        if (cal_on_model) {
          out.x = __dsub_rn(d, __dmul_rn(acc[mi][ni][0], cal0));
          out.y = __dsub_rn(d, __dmul_rn(acc[mi][ni][1], cal1));
        } else {
          out.x = __dsub_rn(__dmul_rn(cal0, d), acc[mi][ni][0]);
          out.y = __dsub_rn(__dmul_rn(cal1, d), acc[mi][ni][1]);
        }
        *reinterpret_cast<double2*>(R + (long long)(row0 + rl) * ldr + col) = out;
      }
    }
  }
}

// ---- the DMMA binning kernel, warp specialised, with Blackwell tile movement and split-ell clusters ------
// Roles inside one CTA (13 warps, one CTA per SM):
//   warps 0..7   MMA consumers: the 64 x 128 tile product, 32 x 32 per warp -- DMMA and shared-memory loads only
//   warps 8..11  BUILDERS: thread = walker; build the [16 ell x 128 walker] Cl chunk from the staged ell-table rows
//                (two independent 8-column recurrences per thread, the arithmetic of k_bin_residual bit for bit)
//   warp 12      PRODUCER: per chunk ONE bulk copy of the tile-major window tile and ONE of the 16 ell-table rows
// Three rings of stages joined by mbarriers (window tiles: producer -> MMA; ell-table rows: producer -> builders;
// Cl chunks: builders -> MMA).  The builders' dependent DFMA / exp chains overlap the MMA warps' DMMA stream by
// construction instead of alternating with it between CTA barriers, and no warp waits for all the others.
//
// Split-ell: a thread-block CLUSTER of S CTAs shares one (window tile, walker tile); CTA r of the cluster takes
// the r-th part of the tile's ell chunks.  The S partial tiles meet in DISTRIBUTED SHARED MEMORY: every CTA parks
// its accumulators in its own shared memory, and CTA r sums rows [64 r / S, 64 (r + 1) / S) of all S partials in
// RANK ORDER (fixed order: bitwise reproducible) and writes their residuals.  S is a property of the PROGRAM (host:
// from the number of chunks per tile), never of the launch size, so a walker's bits do not depend on the batch it
// is evaluated in.  Small launches (47 SPT bandpowers x 4096 chains = 32 tiles) thus fill 4 x as many SMs, and
// 1.3-wave launches become 5-wave launches.
#define BW_MMA_WARPS 8
#define BW_BUILD_WARPS 4
#define BW_THREADS ((BW_MMA_WARPS + BW_BUILD_WARPS + 1) * 32)
#define BW_SA 3
#define BW_SB 3
#define BW_SL 3

template <int NT, int NP>
__global__ void __launch_bounds__(BW_THREADS, 1)
k_bin_residual_ws(csl_program_t p, int tile0, int W, int nsplit, const double* __restrict__ coef, int ldc,
                  double* __restrict__ R, int ldr) {
  constexpr int LSTRIDE = NT + 4 * NP;
  constexpr int LROWS = CSL_KC * LSTRIDE;                    // doubles of one chunk's ell-table rows
  extern __shared__ __align__(128) double smem[];
  double* As = smem;                                         // [BW_SA][CSL_ASZ]
  double* Bs = As + BW_SA * CSL_ASZ;                         // [BW_SB][CSL_BSZ]
  double* Ls = Bs + BW_SB * CSL_BSZ;                         // [BW_SL][LROWS]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ls + BW_SL * LROWS);
  uint64_t *fullA = bars, *emptyA = fullA + BW_SA, *fullB = emptyA + BW_SA, *emptyB = fullB + BW_SB,
           *fullL = emptyB + BW_SB, *emptyL = fullL + BW_SL;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = (int)(blockIdx.x % (unsigned)nsplit);    // == %cluster_ctarank (cluster dims (nsplit, 1, 1))
  const int32_t* tile = p.tile_tab + (long long)(tile0 + blockIdx.x / nsplit) * CSL_TILE_STRIDE;
  const int32_t* spec = p.spec_tab + (long long)tile[CSL_T_SPEC] * CSL_SPEC_STRIDE;
  const int c0 = blockIdx.y * CSL_BN;
  const int l_begin = tile[CSL_T_LBEGIN], l_end = tile[CSL_T_LEND];
  const int nchunk_all = (l_end - l_begin) / CSL_KC;
  // this CTA's share of the chunks: contiguous, the first (nchunk_all % nsplit) shares one longer
  const int cbase = nchunk_all / nsplit, cextra = nchunk_all % nsplit;
  const int cfirst = split * cbase + min(split, cextra);
  const int nchunk = cbase + (split < cextra ? 1 : 0);
  const int lfirst = l_begin + cfirst * CSL_KC;
  if (tid == 0) {
    for (int s = 0; s < BW_SA; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, BW_MMA_WARPS); }
    for (int s = 0; s < BW_SB; ++s) { mbar_init(fullB + s, BW_BUILD_WARPS); mbar_init(emptyB + s, BW_MMA_WARPS); }
    for (int s = 0; s < BW_SL; ++s) { mbar_init(fullL + s, 1); mbar_init(emptyL + s, BW_BUILD_WARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  double acc[4][4][2];
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;        // meaningful for the MMA warps only
  if (warp == BW_MMA_WARPS + BW_BUILD_WARPS) {
    // ---------------------------------------------------------------- producer
    const long long wt_off = ((long long)tile[CSL_T_WINT_HI] << 32) | (unsigned)tile[CSL_T_WINT_LO];
    const double* __restrict__ At = p.winT + wt_off + (long long)(lfirst / CSL_KC) * CSL_ATILE;
    const double* __restrict__ Lt = p.templates + spec[CSL_S_LTAB] + (long long)lfirst * LSTRIDE;
    for (int c = 0; c < nchunk; ++c) {
      if (lane == 0) {
        const int sa = c % BW_SA, sl = c % BW_SL;
        if (c >= BW_SL) mbar_wait(emptyL + sl, ((c / BW_SL) - 1) & 1);
        mbar_arrive_expect_tx(fullL + sl, LROWS * sizeof(double));
        bulk_g2s(Ls + sl * LROWS, Lt + (long long)c * LROWS, LROWS * sizeof(double), fullL + sl);
        if (c >= BW_SA) mbar_wait(emptyA + sa, ((c / BW_SA) - 1) & 1);
        mbar_arrive_expect_tx(fullA + sa, CSL_ASZ * sizeof(double));
        bulk_g2s(As + sa * CSL_ASZ, At + (long long)c * CSL_ATILE, CSL_ASZ * sizeof(double), fullA + sa);
      }
    }
  } else if (warp >= BW_MMA_WARPS) {
    // ---------------------------------------------------------------- builders: thread = walker
    const int wl = tid - BW_MMA_WARPS * 32;
    double cf[NT > 0 ? NT : 1], pamp[NP > 0 ? NP : 1], pidx[NP > 0 ? NP : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) cf[t] = coef[(long long)spec[CSL_S_TERM_COEF + t] * ldc + c0 + wl];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      pamp[q] = coef[(long long)spec[CSL_S_PLAW_AMP + q] * ldc + c0 + wl];
      pidx[q] = coef[(long long)spec[CSL_S_PLAW_IDX + q] * ldc + c0 + wl];
    }
    double bnd[NP > 0 ? NP : 1];
    bool inside[NP > 0 ? NP : 1];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      const float bq = __int_as_float(spec[CSL_S_IDXBOUND + q]);
      bnd[q] = fabs((double)bq);
      inside[q] = bq < 0.f || fabs(pidx[q]) <= bnd[q];
    }
    for (int c = 0; c < nchunk; ++c) {
      const int sl = c % BW_SL, sb = c % BW_SB;
      mbar_wait(fullL + sl, (c / BW_SL) & 1);
      if (c >= BW_SB) mbar_wait(emptyB + sb, ((c / BW_SB) - 1) & 1);
      const double* L = Ls + sl * LROWS;
      double* B = Bs + sb * CSL_BSZ;
      // two independent 8-column recurrences (columns 0..7 and 8..15), interleaved for instruction-level
      // parallelism: per column exactly the arithmetic of k_bin_residual's build_B
      const double* row0 = L;
      const double* row1 = L + (CSL_KC / 2) * LSTRIDE;
      double E0[NP > 0 ? NP : 1], E1[NP > 0 ? NP : 1];
#pragma unroll
      for (int q = 0; q < NP; ++q) {
        E0[q] = exp(pidx[q] * row0[NT + 4 * q]);
        E1[q] = exp(pidx[q] * row1[NT + 4 * q]);
      }
#pragma unroll 2
      for (int r = 0; r < CSL_KC / 2; ++r, row0 += LSTRIDE, row1 += LSTRIDE) {
        double v0 = 0.0, v1 = 0.0;
#pragma unroll
        for (int t = 0; t < NT; ++t) { v0 = fma(cf[t], row0[t], v0); v1 = fma(cf[t], row1[t], v1); }
#pragma unroll
        for (int q = 0; q < NP; ++q) {
          v0 = fma(pamp[q] * row0[NT + 4 * q + 1], E0[q], v0);
          v1 = fma(pamp[q] * row1[NT + 4 * q + 1], E1[q], v1);
        }
        B[r * CSL_LDB + wl] = v0;
        B[(CSL_KC / 2 + r) * CSL_LDB + wl] = v1;
        if (r + 1 < CSL_KC / 2) {
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const double D0 = row0[NT + 4 * q + 2], D1 = row1[NT + 4 * q + 2];
            const double y0 = pidx[q] * D0, ay0 = bnd[q] * fabs(D0);      // ay: warp-uniform (broadcast row)
            if (ay0 <= 0.012 && inside[q]) E0[q] *= exp_taylor<6>(y0);
            else if (ay0 <= 0.1 && inside[q]) E0[q] *= exp_taylor<10>(y0);
            else E0[q] = exp(pidx[q] * row0[LSTRIDE + NT + 4 * q]);
            const double y1 = pidx[q] * D1, ay1 = bnd[q] * fabs(D1);
            if (ay1 <= 0.012 && inside[q]) E1[q] *= exp_taylor<6>(y1);
            else if (ay1 <= 0.1 && inside[q]) E1[q] *= exp_taylor<10>(y1);
            else E1[q] = exp(pidx[q] * row1[LSTRIDE + NT + 4 * q]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) { mbar_arrive(fullB + sb); mbar_arrive(emptyL + sl); }
    }
  } else {
    // ---------------------------------------------------------------- MMA consumers
    int glo[4], ghi[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      glo[mi] = tile[CSL_T_GLO + (wm0 >> 3) + mi];
      ghi[mi] = tile[CSL_T_GHI + (wm0 >> 3) + mi];
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      const int sa = c % BW_SA, sb = c % BW_SB;
      const int l0 = lfirst + c * CSL_KC;
      unsigned mmask = 0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
        if (l0 + CSL_KC > glo[mi] && l0 < ghi[mi]) mmask |= 1u << mi;
      mbar_wait(fullA + sa, (c / BW_SA) & 1);
      mbar_wait(fullB + sb, (c / BW_SB) & 1);
      if (mmask) warp_mma_chunk(acc, As + sa * CSL_ASZ, Bs + sb * CSL_BSZ, wm0, wn0, lane, mmask);
      __syncwarp();
      if (lane == 0) { mbar_arrive(emptyA + sa); mbar_arrive(emptyB + sb); }
    }
  }

  const int g = lane >> 2, t = lane & 3;
  const int row0 = tile[CSL_T_ROW0], nrows = tile[CSL_T_NROWS];
  const int calrow = spec[CSL_S_CAL];
  const bool cal_on_model = spec[CSL_S_CALMODE] != 0;      // r = d - cal*b  (SPTSZ_lowl.py:68)
  if (nsplit == 1) {
    if (warp >= BW_MMA_WARPS) return;
    // epilogue: r = cal * d - b   (spt_lowl.py:132), straight from the accumulators
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int col = c0 + wn0 + ni * 8 + 2 * t;
      double cal0 = 1.0, cal1 = 1.0;
      if (calrow >= 0) {
        cal0 = coef[(long long)calrow * ldc + col];
        cal1 = coef[(long long)calrow * ldc + col + 1];
      }
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int rl = wm0 + mi * 8 + g;
        if (rl < nrows) {
          const double d = p.data[row0 + rl];
          double2 out;
          if (cal_on_model) {
            out.x = __dsub_rn(d, __dmul_rn(acc[mi][ni][0], cal0));
            out.y = __dsub_rn(d, __dmul_rn(acc[mi][ni][1], cal1));
          } else {
            out.x = __dsub_rn(__dmul_rn(cal0, d), acc[mi][ni][0]);
            out.y = __dsub_rn(__dmul_rn(cal1, d), acc[mi][ni][1]);
          }
          *reinterpret_cast<double2*>(R + (long long)(row0 + rl) * ldr + col) = out;
        }
      }
    }
    return;
  }
  // ---- split-ell: park the partial tile in shared memory (the rings are idle now), meet in DSMEM
  double* X = smem;                                         // [64][BW_XLD], over the A / B rings
  if (warp < BW_MMA_WARPS) {
    asm volatile("bar.sync 1, %0;" :: "n"(BW_MMA_WARPS * 32) : "memory");     // every MMA warp has left the rings
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        double2 v;
        v.x = acc[mi][ni][0];
        v.y = acc[mi][ni][1];
        *reinterpret_cast<double2*>(X + (wm0 + mi * 8 + g) * BW_XLD + wn0 + ni * 8 + 2 * t) = v;
      }
  }
  cluster_sync_all();                                       // all threads of all CTAs of the cluster
  if (warp < BW_MMA_WARPS) {
    const int rows_per = CSL_BM / nsplit;                   // nsplit in {2, 4}
    const int r_lo = split * rows_per;
    for (int e = tid; e < rows_per * CSL_BN; e += BW_MMA_WARPS * 32) {
      const int rl = r_lo + e / CSL_BN, cl = e % CSL_BN;
      const double* src = X + rl * BW_XLD + cl;
      double b = ld_dsmem(src, 0);
      for (int q = 1; q < nsplit; ++q) b += ld_dsmem(src, (uint32_t)q);      // rank order: fixed
      if (rl < nrows) {
        const int col = c0 + cl;
        const double cal = calrow >= 0 ? coef[(long long)calrow * ldc + col] : 1.0;
        const double d = p.data[row0 + rl];
        R[(long long)(row0 + rl) * ldr + col] = cal_on_model ? __dsub_rn(d, __dmul_rn(b, cal)) : __dsub_rn(__dmul_rn(cal, d), b);
      }
    }
  }
  cluster_sync_all();                                       // nobody leaves while a peer still reads its tile
}

template <int NT, int NP>
static int launch_class_ws(const csl_program_t* p, int tile0, int ntile, int nsplit, int W, const double* coef, int ldc,
                           double* R, int ldr, cudaStream_t s) {
  const size_t smem = (size_t)(BW_SA * CSL_ASZ + BW_SB * CSL_BSZ + BW_SL * CSL_KC * (NT + 4 * NP)) * sizeof(double) +
                      2 * (BW_SA + BW_SB + BW_SL) * sizeof(uint64_t);
  static_assert(BW_SA * CSL_ASZ + BW_SB * CSL_BSZ >= CSL_BM * BW_XLD, "the rings must hold the parked partial tile");
  cudaError_t e = cudaFuncSetAttribute(k_bin_residual_ws<NT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ntile * nsplit, W / CSL_BN, 1);
  cfg.blockDim = dim3(BW_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nsplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, k_bin_residual_ws<NT, NP>, *p, tile0, W, nsplit, coef, ldc, R, ldr);
  if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  return 0;
}

#define CSL_DISPATCH_WS_NP(NTV)                                                                                  \
  switch (np) {                                                                                                  \
    case 0: return launch_class_ws<NTV, 0>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);                    \
    case 1: return launch_class_ws<NTV, 1>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);                    \
    case 2: return launch_class_ws<NTV, 2>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);                    \
    case 4: return launch_class_ws<NTV, 4>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);                    \
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: power-law class must be 0, 1, 2 or 4");            \
  }

static int dispatch_class_ws(int nt, int np, const csl_program_t* p, int tile0, int ntile, int nsplit, int W,
                             const double* coef, int ldc, double* R, int ldr, cudaStream_t s) {
  switch (nt) {
    case 0: CSL_DISPATCH_WS_NP(0)
    case 2: CSL_DISPATCH_WS_NP(2)
    case 4: CSL_DISPATCH_WS_NP(4)
    case 8: CSL_DISPATCH_WS_NP(8)
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: template class must be 0, 2, 4 or 8");
  }
}

// ---- segmented-sum binning for block-sparse windows ----------------------------------------------------
// One thread per walker (SEG_WPT walkers per thread), CSL_SEG_GROUP window rows per CTA.  For each row only
// its non-zero support [lo, hi) is visited.  The columns of the group's rows lie back to back in a table the
// host builds once (after the ell-tables in `templates`, bin record word CSL_B_TABOFF), ALREADY MULTIPLIED by
// the window weight -- the weight is the same for every walker -- so staging is one contiguous cp.async copy:
//     staged row = [w T_0 .. w T_{NT-1},  w M_0 .. w M_{NP-1},  D_0 .. D_{NP-1}]          (RS doubles)
// and the inner loop is broadcast LDS + FP64 FMAs only:
//     b += sum_t coef_t (w T_t) + sum_q (w M_q) Ehat_q ,      Ehat_q = amp_q exp(idx_q logB_q[l]).
//
// Power laws along consecutive l:  E[l] = exp(idx logB[l]) obeys E[l+1] = E[l] exp(idx D[l]) with
// D[l] = logB[l+1] - logB[l] (host, extended precision).  Within a bin the kernel takes one exact exp at
// the first column (times the amplitude) and advances by a Taylor polynomial in y = idx D: degree 5 when
// |y| <= 0.0035 (truncation y^6/6! < 3e-18), 6 when |y| <= 0.012 (< 7e-18), 10 when |y| <= 0.1 (< 3e-19),
// decided ONCE per group from the largest |D| of its bins (bin records) times the host's bound on |idx|
// (spectrum record) -- otherwise, and for walkers outside the bound, exp at every column.
// A bin that begins at the column where the previous bin of the group ended (top-hat bins) inherits the
// state instead of taking a new exp.  Rounding grows by ~1 ulp per column (<= 8 bins x a few tens of
// columns: a few 1e-14 relative at worst), far inside the 1e-10 budget on lnL.
//
// A group whose columns exceed the staging capacity is processed in several passes; the bins are walked in
// order, so a bin cut by a pass boundary simply keeps its running sum and power-law state in registers.
#ifndef SEG_TPB
#define SEG_TPB 128    // threads per CTA
#endif
// WPT = walkers per thread (template parameter): 2 in bulk -- staged rows / loop control are amortised over both --
// and 1 for launches of a few thousand walkers, where twice as many CTAs and warps hide latency better
#ifndef SEG_SMEM_DOUBLES
#define SEG_SMEM_DOUBLES 3072   // staged columns x RS (24 KB)
#endif
#ifndef SEG_MINB
#define SEG_MINB 1
#endif
#ifndef SEG_UNROLL
#define SEG_UNROLL 2
#endif

// staged-row stride: NT real template terms + NP (w M_q) + NP D_q, padded to a whole number of 16-byte pieces
__host__ __device__ constexpr int seg_row_stride(int nt, int np) { return (nt + 2 * np + 1) & ~1; }
// the padded template-term class {0, 2, 4, 8} the ell-tables are laid out for (k_bin_residual's classes)
__host__ __device__ constexpr int seg_pad_terms(int nt) { return nt <= 0 ? 0 : nt <= 2 ? 2 : nt <= 4 ? 4 : 8; }

// ncol staged columns starting at `row`, for the SEG_WPT walkers of a thread.
// MODE 0 / 1 / 2: Taylor degree 5 / 6 / 10 in y = idx D;  MODE 3: exact exp at every column (ltab, l = ell
// index of the first column, give the next column's logB; lend = one past the bin's last column).
template <int NT, int NP, int MODE, int SEG_WPT>
__device__ __forceinline__ void seg_cols(const double* __restrict__ row, int ncol,
                                         const double (&cf)[SEG_WPT][NT > 0 ? NT : 1],
                                         const double (&pamp)[SEG_WPT][NP > 0 ? NP : 1],
                                         const double (&pidx)[SEG_WPT][NP > 0 ? NP : 1],
                                         double (&E)[SEG_WPT][NP > 0 ? NP : 1], double (&at)[SEG_WPT],
                                         double (&ap)[SEG_WPT], const double* __restrict__ ltab, int l, int lend) {
  // NT is the spectrum's REAL number of template terms; its ell-table rows are laid out for the padded class
  constexpr int RS = seg_row_stride(NT, NP), NTP = seg_pad_terms(NT), LSTRIDE = NTP + 4 * NP;
  auto column = [&](const double* __restrict__ r, int lcur) {
    double T[NT > 0 ? NT : 1], M[NP > 0 ? NP : 1], D[NP > 0 ? NP : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) T[t] = r[t];
#pragma unroll
    for (int q = 0; q < NP; ++q) { M[q] = r[NT + q]; D[q] = r[NT + NP + q]; }
#pragma unroll
    for (int j = 0; j < SEG_WPT; ++j) {
#pragma unroll
      for (int t = 0; t < NT; ++t) at[j] = fma(cf[j][t], T[t], at[j]);
#pragma unroll
      for (int q = 0; q < NP; ++q) ap[j] = fma(M[q], E[j][q], ap[j]);
    }
#pragma unroll
    for (int q = 0; q < NP; ++q) {
#pragma unroll
      for (int j = 0; j < SEG_WPT; ++j) {
        if (MODE == 0) E[j][q] *= exp_taylor<5>(pidx[j][q] * D[q]);
        else if (MODE == 1) E[j][q] *= exp_taylor<6>(pidx[j][q] * D[q]);
        else if (MODE == 2) E[j][q] *= exp_taylor<10>(pidx[j][q] * D[q]);
        else if (lcur + 1 < lend)
          E[j][q] = pamp[j][q] * exp(pidx[j][q] * __ldg(ltab + (long long)(lcur + 1) * LSTRIDE + NTP + 4 * q));
      }
    }
  };
  int c = 0;
  for (; c + SEG_UNROLL <= ncol; c += SEG_UNROLL, row += SEG_UNROLL * RS) {
#pragma unroll
    for (int u = 0; u < SEG_UNROLL; ++u) column(row + u * RS, l + c + u);
  }
#pragma unroll 1
  for (; c < ncol; ++c, row += RS) column(row, l + c);
}

// one CTA's work: group `gy` of the class that starts at bin `bin0`, SEG_WPT * SEG_TPB walkers from blockIdx.x on;
// srow / sg are the CTA's staging area and group record (shared memory, owned by the kernel)
template <int NT, int NP, int SEG_WPT>
__device__ __forceinline__ void seg_group(const csl_program_t& p, int bin0, int gy, int W, const double* __restrict__ coef,
                                          int ldc, double* __restrict__ R, int ldr, double* srow, int32_t* sg) {
  constexpr int RS = seg_row_stride(NT, NP);         // > 0: the host never emits an empty class
  constexpr int CAP = SEG_SMEM_DOUBLES / (RS > 0 ? RS : 1);
  constexpr int NPX = NP > 0 ? NP : 1, NTX = NT > 0 ? NT : 1;
  const int tid = threadIdx.x;
  int w[SEG_WPT];
  bool live_w[SEG_WPT];
#pragma unroll
  for (int j = 0; j < SEG_WPT; ++j) {
    w[j] = blockIdx.x * (SEG_WPT * SEG_TPB) + j * SEG_TPB + tid;
    live_w[j] = w[j] < W;
    if (!live_w[j]) w[j] = W - 1;             // replay a valid walker, never stored
  }
  // bin records come in whole groups and classes start on a group boundary: group = bin0 / 8 + blockIdx.y
  const int32_t* __restrict__ grec = p.grp_tab + (long long)(bin0 / CSL_SEG_GROUP + gy) * CSL_GRP_STRIDE;
  if (tid < CSL_GRP_STRIDE / 4)
    reinterpret_cast<int4*>(sg)[tid] = __ldg(reinterpret_cast<const int4*>(grec) + tid);
  const int32_t* spec = p.spec_tab + (long long)__ldg(grec + CSL_G_SPEC) * CSL_SPEC_STRIDE;   // one spectrum per group
  const double* __restrict__ ltab = p.templates + spec[CSL_S_LTAB];
  // columns p0 .. p0+CAP of the group -> shared memory (weighted rows, contiguous in the host-built table).
  // The first pass is issued here, from the record in global memory, so it flies together with the record
  // and coefficient loads; one barrier later everything is in place.
  const int total = __ldg(grec + CSL_G_PRE + CSL_SEG_GROUP);
  const long long sbase = ((long long)__ldg(grec + CSL_G_BASE_HI) << 32) | (unsigned)__ldg(grec + CSL_G_BASE_LO);
  auto stage = [&](int p0) {
    const double* __restrict__ gsrc = p.templates + sbase + (long long)p0 * RS;
    const int n16 = (min(p0 + CAP, total) - p0) * RS / 2;
    for (int i = tid; i < n16; i += SEG_TPB) cp_async16(srow + 2 * i, gsrc + 2 * i);
    cp_async_commit();
  };
  stage(0);
  const int* spre = sg + CSL_G_PRE;
  const int* slo = sg + CSL_G_LO;
  const int* srw = sg + CSL_G_ROW;
  const int* scont = sg + CSL_G_CONT;
  const double* sdat = reinterpret_cast<const double*>(sg + CSL_G_DATA);
  const double (*slogb)[CSL_MAX_PLAW] = reinterpret_cast<const double (*)[CSL_MAX_PLAW]>(sg + CSL_G_LOGB);
  double cf[SEG_WPT][NTX], pamp[SEG_WPT][NPX], pidx[SEG_WPT][NPX], cal[SEG_WPT];
  const int calrow = spec[CSL_S_CAL];
  const bool cal_on_model = spec[CSL_S_CALMODE] != 0;      // r = d - cal*b  (SPTSZ_lowl.py:68)
#pragma unroll
  for (int j = 0; j < SEG_WPT; ++j) {
#pragma unroll
    for (int t = 0; t < NT; ++t) cf[j][t] = coef[(long long)spec[CSL_S_TERM_COEF + t] * ldc + w[j]];
#pragma unroll
    for (int q = 0; q < NP; ++q) {
      pamp[j][q] = coef[(long long)spec[CSL_S_PLAW_AMP + q] * ldc + w[j]];
      pidx[j][q] = coef[(long long)spec[CSL_S_PLAW_IDX + q] * ldc + w[j]];
    }
    cal[j] = calrow >= 0 ? coef[(long long)calrow * ldc + w[j]] : 1.0;
  }
  cp_async_wait<0>();
  __syncthreads();
  double E[SEG_WPT][NPX], at[SEG_WPT], ap[SEG_WPT];       // running state of the bin being summed
#pragma unroll
  for (int j = 0; j < SEG_WPT; ++j) {
    at[j] = ap[j] = 0.0;
#pragma unroll
    for (int q = 0; q < NPX; ++q) E[j][q] = 0.0;
  }
  // Taylor degree: ONE static choice per group (group record, computed by the host from the largest |D| of
  // its bins and its bound on |idx|: the prior box of the index parameter, else 4) -- the same
  // for every walker, so a walker's arithmetic never depends on which walkers share its thread, warp or GPU.
  // A walker whose index lies outside a SOFT bound takes the exact-exp mode instead; outside a hard bound (a
  // prior box, stored negative) its lnl is +inf whatever this kernel computes, so it needs no special care.
  const int smode = sg[CSL_G_MODE];
  int mode[SEG_WPT];
#pragma unroll
  for (int j = 0; j < SEG_WPT; ++j) {
    bool inside = true;
#pragma unroll
    for (int q = 0; q < NP; ++q) {   // hard bounds (stored negative) need no check; NaN -> outside
      const float bq = __int_as_float(spec[CSL_S_IDXBOUND + q]);
      inside = inside && (bq < 0.f || fabs(pidx[j][q]) <= (double)bq);
    }
    mode[j] = inside ? smode : 3;
  }
  // if any thread of the CTA holds two walkers of different modes, every thread runs the group twice, once
  // per walker slot at that walker's own mode (only the slot's results are stored): still one code path
  const int split = SEG_WPT > 1 ? __syncthreads_or(mode[0] != mode[SEG_WPT - 1]) : 0;

  for (int run = 0; run < (split ? SEG_WPT : 1); ++run) {
  const int run_mode = split ? mode[run] : mode[0];
  const unsigned store_mask = split ? (1u << run) : ~0u;
  for (int p0 = 0; p0 < total; p0 += CAP) {
    const int p1 = min(p0 + CAP, total);
    if (p0 > 0 || (run > 0 && total > CAP)) {        // the first pass of the first run is already staged
      stage(p0);
      cp_async_wait<0>();
      __syncthreads();
    }
    auto run_bins = [&](auto tag) {
      constexpr int MODE = decltype(tag)::value;
#pragma unroll 1
      for (int b = 0; b < CSL_SEG_GROUP; ++b) {
        const int s0 = max(spre[b], p0), s1 = min(spre[b + 1], p1);
        if (s1 <= s0) continue;                   // block-uniform
        const int l = slo[b] + (s0 - spre[b]);
        if (s0 == spre[b]) {                      // the bin starts here
#pragma unroll
          for (int j = 0; j < SEG_WPT; ++j) at[j] = ap[j] = 0.0;
          // a bin that begins where the previous one ended inherits its power-law state: the update
          // after a bin's last column has already advanced E to this column
          if (MODE == 3 || !scont[b]) {
#pragma unroll
            for (int j = 0; j < SEG_WPT; ++j)
#pragma unroll
              for (int q = 0; q < NP; ++q) E[j][q] = pamp[j][q] * exp(pidx[j][q] * slogb[b][q]);
          }
        }
        const int lend = slo[b] + (spre[b + 1] - spre[b]);
        seg_cols<NT, NP, MODE, SEG_WPT>(srow + (s0 - p0) * RS, s1 - s0, cf, pamp, pidx, E, at, ap, ltab, l, lend);
        if (s1 == spre[b + 1]) {                  // the bin ends here: r = cal * d - b  (spt_lowl.py:132)
          const int row = srw[b];
          const double d = sdat[b];
#pragma unroll
          for (int j = 0; j < SEG_WPT; ++j) {
            const double tot = at[j] + ap[j];
            if (live_w[j] && (store_mask >> j & 1u))
              R[(long long)row * ldr + w[j]] = cal_on_model ? __dsub_rn(d, __dmul_rn(tot, cal[j]))
                                                            : __dsub_rn(__dmul_rn(cal[j], d), tot);
          }
        }
      }
    };
    switch (run_mode) {
      case 0: run_bins(std::integral_constant<int, 0>{}); break;
      case 1: run_bins(std::integral_constant<int, 1>{}); break;
      case 2: run_bins(std::integral_constant<int, 2>{}); break;
      default: run_bins(std::integral_constant<int, 3>{}); break;
    }
    __syncthreads();
  }
  }   // run
  // rows whose window is empty (no columns at all) still owe r = cal * d
#pragma unroll 1
  for (int b = 0; b < CSL_SEG_GROUP; ++b) {
    const int row = srw[b];
    if (row >= 0 && spre[b + 1] == spre[b]) {
#pragma unroll
      for (int j = 0; j < SEG_WPT; ++j)
        if (live_w[j]) R[(long long)row * ldr + w[j]] = cal_on_model ? sdat[b] : __dmul_rn(cal[j], sdat[b]);
    }
  }
}

template <int NT, int NP, int SEG_WPT>
__global__ void __launch_bounds__(SEG_TPB, SEG_MINB)
k_bin_segsum(csl_program_t p, int bin0, int nbins, int W, const double* __restrict__ coef, int ldc,
             double* __restrict__ R, int ldr) {
  __shared__ __align__(16) double srow[SEG_SMEM_DOUBLES];
  __shared__ __align__(16) int32_t sg[CSL_GRP_STRIDE];          // the group record (host-built, CSL_G_*)
  seg_group<NT, NP, SEG_WPT>(p, bin0, (int)blockIdx.y, W, coef, ldc, R, ldr, srow, sg);
}

// TWO spectrum classes in one launch (grid.y = groups of class A, then groups of class B): the classes write
// different rows of R and read the same coefficients, so nothing orders them; as two launches the second waited
// for the first one's last CTAs (a few us per half-step, which is several per cent of a 4096-walker launch).
// Instantiated for the class pairs that occur (launch_seg_pair); any other pair takes the two launches.
template <int NTA, int NPA, int NTB, int NPB, int SEG_WPT>
__global__ void __launch_bounds__(SEG_TPB, 5)          // <= 96 registers: five CTAs per SM like the single-class kernels (98 gave four)
k_bin_segsum_pair(csl_program_t p, int bin0a, int ngrpa, int bin0b, int W, const double* __restrict__ coef, int ldc,
                  double* __restrict__ R, int ldr) {
  __shared__ __align__(16) double srow[SEG_SMEM_DOUBLES];
  __shared__ __align__(16) int32_t sg[CSL_GRP_STRIDE];
  if ((int)blockIdx.y < ngrpa) seg_group<NTA, NPA, SEG_WPT>(p, bin0a, (int)blockIdx.y, W, coef, ldc, R, ldr, srow, sg);
  else seg_group<NTB, NPB, SEG_WPT>(p, bin0b, (int)blockIdx.y - ngrpa, W, coef, ldc, R, ldr, srow, sg);
}

static int seg_small_limit() {
  int small = csl_option_value("seg_small");
  return small < 0 ? 8192 : small;
}

template <int NTA, int NPA, int NTB, int NPB>
static int launch_seg_pair(const csl_program_t* p, const int32_t* ka, const int32_t* kb, int W, const double* coef,
                           int ldc, double* R, int ldr, cudaStream_t s) {
  const int nga = (ka[3] + CSL_SEG_GROUP - 1) / CSL_SEG_GROUP, ngb = (kb[3] + CSL_SEG_GROUP - 1) / CSL_SEG_GROUP;
  if (W <= seg_small_limit()) {
    dim3 grid((W + SEG_TPB - 1) / SEG_TPB, nga + ngb);
    k_bin_segsum_pair<NTA, NPA, NTB, NPB, 1><<<grid, SEG_TPB, 0, s>>>(*p, ka[2], nga, kb[2], W, coef, ldc, R, ldr);
  } else {
    dim3 grid((W + 2 * SEG_TPB - 1) / (2 * SEG_TPB), nga + ngb);
    k_bin_segsum_pair<NTA, NPA, NTB, NPB, 2><<<grid, SEG_TPB, 0, s>>>(*p, ka[2], nga, kb[2], W, coef, ldc, R, ldr);
  }
  return 0;
}

// the class pairs with a fused launch: (template terms, power laws) of the first and the second class
static bool dispatch_seg_pair(const csl_program_t* p, const int32_t* ka, const int32_t* kb, int W, const double* coef,
                              int ldc, double* R, int ldr, cudaStream_t s) {
  if (csl_option_value("seg_pair") == 0) return false;      // option / CSL_SEG_PAIR=0: one launch per class
#define CSL_SEG_PAIR(A, B, C, D)                                                                  \
  if (ka[0] == A && ka[1] == B && kb[0] == C && kb[1] == D) {                                     \
    launch_seg_pair<A, B, C, D>(p, ka, kb, W, coef, ldc, R, ldr, s);                              \
    return true;                                                                                  \
  }
  CSL_SEG_PAIR(3, 2, 1, 1)      // TT with two power laws + TE / EE with one (plik-lite-like TT/TE/EE)
  CSL_SEG_PAIR(1, 1, 3, 2)
  CSL_SEG_PAIR(2, 1, 1, 1)
  CSL_SEG_PAIR(1, 1, 2, 1)
  CSL_SEG_PAIR(4, 2, 2, 1)
  CSL_SEG_PAIR(2, 1, 4, 2)
#undef CSL_SEG_PAIR
  return false;
}

template <int NT, int NP>
static int launch_seg(const csl_program_t* p, int bin0, int nbins, int W, const double* coef, int ldc,
                      double* R, int ldr, cudaStream_t s) {
  // one walker per thread for launches of a few thousand walkers ("seg_small": the largest such launch, default
  // 8192): the same arithmetic per walker, twice the CTAs
  int small = csl_option_value("seg_small");
  if (small < 0) small = 8192;
  const int ngrp = (nbins + CSL_SEG_GROUP - 1) / CSL_SEG_GROUP;
  if (W <= small) {
    dim3 grid((W + SEG_TPB - 1) / SEG_TPB, ngrp);
    k_bin_segsum<NT, NP, 1><<<grid, SEG_TPB, 0, s>>>(*p, bin0, nbins, W, coef, ldc, R, ldr);
  } else {
    dim3 grid((W + 2 * SEG_TPB - 1) / (2 * SEG_TPB), ngrp);
    k_bin_segsum<NT, NP, 2><<<grid, SEG_TPB, 0, s>>>(*p, bin0, nbins, W, coef, ldc, R, ldr);
  }
  return 0;
}

#define CSL_DISPATCH_SEG_NP(NTV)                                                                       \
  switch (np) {                                                                                        \
    case 0: return launch_seg<NTV, 0>(p, b0, nb, W, coef, ldc, R, ldr, s);                             \
    case 1: return launch_seg<NTV, 1>(p, b0, nb, W, coef, ldc, R, ldr, s);                             \
    case 2: return launch_seg<NTV, 2>(p, b0, nb, W, coef, ldc, R, ldr, s);                             \
    case 4: return launch_seg<NTV, 4>(p, b0, nb, W, coef, ldc, R, ldr, s);                             \
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: power-law class must be 0, 1, 2 or 4");  \
  }

static int dispatch_seg(int nt, int np, const csl_program_t* p, int b0, int nb, int W, const double* coef, int ldc,
                        double* R, int ldr, cudaStream_t s) {
  switch (nt) {           // the REAL number of template terms: no FMA is spent on padding terms
    case 0: CSL_DISPATCH_SEG_NP(0)
    case 1: CSL_DISPATCH_SEG_NP(1)
    case 2: CSL_DISPATCH_SEG_NP(2)
    case 3: CSL_DISPATCH_SEG_NP(3)
    case 4: CSL_DISPATCH_SEG_NP(4)
    case 5: CSL_DISPATCH_SEG_NP(5)
    case 6: CSL_DISPATCH_SEG_NP(6)
    case 7: CSL_DISPATCH_SEG_NP(7)
    case 8: CSL_DISPATCH_SEG_NP(8)
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: template terms must be 0..8");
  }
}

template <int NT, int NP>
static int launch_class(const csl_program_t* p, int tile0, int ntile, int nsplit, int W, const double* coef, int ldc,
                        double* R, int ldr, cudaStream_t s) {
  size_t smem = (size_t)(2 * CSL_ASZ + 2 * CSL_BSZ + 2 * CSL_KC * (NT + 4 * NP)) * sizeof(double);
  if (nsplit > 1 && smem < (size_t)CSL_BM * BW_XLD * sizeof(double)) smem = (size_t)CSL_BM * BW_XLD * sizeof(double);
  {
    cudaError_t e = cudaFuncSetAttribute(k_bin_residual<NT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  if (nsplit <= 1) {
    dim3 grid(ntile, W / CSL_BN);
    k_bin_residual<NT, NP><<<grid, CSL_THREADS, smem, s>>>(*p, tile0, W, 1, coef, ldc, R, ldr);
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ntile * nsplit, W / CSL_BN, 1);
  cfg.blockDim = dim3(CSL_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nsplit;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_bin_residual<NT, NP>, *p, tile0, W, nsplit, coef, ldc, R, ldr);
  if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  return 0;
}

#define CSL_DISPATCH_NP(NTV)                                                                          \
  switch (np) {                                                                                        \
    case 0: return launch_class<NTV, 0>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);             \
    case 1: return launch_class<NTV, 1>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);             \
    case 2: return launch_class<NTV, 2>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);             \
    case 4: return launch_class<NTV, 4>(p, tile0, ntile, nsplit, W, coef, ldc, R, ldr, s);             \
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: power-law class must be 0, 1, 2 or 4");  \
  }

static int dispatch_class(int nt, int np, const csl_program_t* p, int tile0, int ntile, int nsplit, int W,
                          const double* coef, int ldc, double* R, int ldr, cudaStream_t s) {
  switch (nt) {
    case 0: CSL_DISPATCH_NP(0)
    case 2: CSL_DISPATCH_NP(2)
    case 4: CSL_DISPATCH_NP(4)
    case 8: CSL_DISPATCH_NP(8)
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: template class must be 0, 2, 4 or 8");
  }
}

// kernel launches one csl_bin_residual call makes for this program (bench.py counts its launches with it)
extern "C" int csl_bin_residual_launches(const csl_program_t* p) {
  if (!p) return 0;
  int n = p->nclass;
  const bool off = csl_option_value("seg_pair") == 0;
  for (int c = 0; c < p->nsegclass; ++c) {
    ++n;
    if (!off && c + 1 < p->nsegclass) {
      const int32_t* a = p->segcls[c];
      const int32_t* b = p->segcls[c + 1];
      const int key[4] = {a[0], a[1], b[0], b[1]};
      static const int pairs[6][4] = {{3, 2, 1, 1}, {1, 1, 3, 2}, {2, 1, 1, 1}, {1, 1, 2, 1}, {4, 2, 2, 1}, {2, 1, 4, 2}};
      for (const auto& q : pairs)
        if (q[0] == key[0] && q[1] == key[1] && q[2] == key[2] && q[3] == key[3]) { ++c; break; }
    }
  }
  return n;
}

extern "C" int csl_bin_residual(const csl_program_t* p, int W, const double* coef, int ldc, double* R, int ldr,
                                void* stream) {
  CSL_REQUIRE(p && coef && R, "csl_bin_residual: null argument");
  CSL_REQUIRE((p->resid_mode == 1 || p->resid_mode == 3) && (p->nclass > 0 || p->nsegclass > 0) &&
                  p->nclass <= CSL_MAX_CLASS &&
                  p->nsegclass <= CSL_MAX_CLASS,
              "csl_bin_residual: program has no binned spectra");
  CSL_REQUIRE(W > 0 && W % CSL_BN == 0 && ldr >= W && ldc >= W && (ldr % 2) == 0,
              "csl_bin_residual: W must be a positive multiple of 128 and ldr, ldc >= W");
  cudaStream_t s = (cudaStream_t)stream;
  // rows n .. n_pad are written (as zeros) by the kernels too: the host appends records with an empty
  // window and zero data for them (program.py), so no separate clearing pass is needed
  for (int c = 0; c < p->nclass; ++c) {
    const int32_t* k = p->cls[c];
    CSL_REQUIRE(k[0] + 2 * k[1] > 0, "csl_bin_residual: empty spectrum class");
    // The program's ell split (csl_program_t::bin_split, a property of the program, never of the launch) becomes
    // the cluster size.  bin_variant 2 selects the warp-specialised bulk-copy kernel (measured SLOWER than the
    // all-warps-build-then-multiply kernel at every size but tiny launches: DESIGN.md); both give the same bits.
    int nsplit = p->bin_split;
    if (csl_option_value("bin_split") > 0) nsplit = csl_option_value("bin_split");       // sweeps only: changes the bits
    if (nsplit != 2 && nsplit != 4) nsplit = 1;
    int rc;
    if (p->winT && csl_option_value("bin_variant") == 2 && !p->has_aberr)
      rc = dispatch_class_ws(k[0], k[1], p, k[2], k[3], nsplit, W, coef, ldc, R, ldr, s);
    else
      rc = dispatch_class(k[0], k[1], p, k[2], k[3], nsplit, W, coef, ldc, R, ldr, s);
    if (rc) return rc;
  }
  for (int c = 0; c < p->nsegclass; ++c) {
    const int32_t* k = p->segcls[c];
    CSL_REQUIRE(k[0] + 2 * k[1] > 0 && p->bin_tab && p->grp_tab, "csl_bin_residual: empty segmented class");
    if (c + 1 < p->nsegclass && dispatch_seg_pair(p, k, p->segcls[c + 1], W, coef, ldc, R, ldr, s)) {
      ++c;                       // both classes went out in one launch
      continue;
    }
    int rc = dispatch_seg(k[0], k[1], p, k[2], k[3], W, coef, ldc, R, ldr, s);
    if (rc) return rc;
  }
  return csl_check_launch("k_bin_residual");
}
