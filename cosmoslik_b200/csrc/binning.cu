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

template <int NT, int NP>
__global__ void __launch_bounds__(CSL_THREADS, 2)
k_bin_residual(csl_program_t p, int tile0, int W, const double* __restrict__ coef, int ldc,
               double* __restrict__ R, int ldr) {
  constexpr int LSTRIDE = NT + 4 * NP;
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                       // [2][CSL_ASZ]
  double* Bs = smem + 2 * CSL_ASZ;         // [2][CSL_BSZ]
  double* Ls = Bs + 2 * CSL_BSZ;           // [2][CSL_KC * LSTRIDE]
  const int32_t* tile = p.tile_tab + (long long)(tile0 + blockIdx.x) * CSL_TILE_STRIDE;
  const int32_t* spec = p.spec_tab + (long long)tile[CSL_T_SPEC] * CSL_SPEC_STRIDE;
  const int c0 = blockIdx.y * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;

  const int l_begin = tile[CSL_T_LBEGIN], l_end = tile[CSL_T_LEND];
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
  auto build_B = [&](double* B, const double* L) {
    const double* row = L + grp * (CSL_KC / 2) * LSTRIDE;     // warp-uniform address -> shared-memory broadcast
    double E[NP > 0 ? NP : 1];
#pragma unroll
    for (int q = 0; q < NP; ++q) E[q] = exp(pidx[q] * row[NT + 4 * q]);
#pragma unroll 2
    for (int r = 0; r < CSL_KC / 2; ++r, row += LSTRIDE) {
      double v = 0.0;
#pragma unroll
      for (int t = 0; t < NT; ++t) v = fma(cf[t], row[t], v);
#pragma unroll
      for (int q = 0; q < NP; ++q) v = fma(pamp[q] * row[NT + 4 * q + 1], E[q], v);
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
    build_B(Bs, Ls);
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
      if (c + 1 < nchunk) build_B(Bs + (cur ^ 1) * CSL_BSZ, Ls + (cur ^ 1) * CSL_KC * LSTRIDE);
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
#define SEG_WPT 2      // walkers per thread: staged rows / loop control are amortised over both
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
template <int NT, int NP, int MODE>
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

template <int NT, int NP>
__global__ void __launch_bounds__(SEG_TPB, SEG_MINB)
k_bin_segsum(csl_program_t p, int bin0, int nbins, int W, const double* __restrict__ coef, int ldc,
             double* __restrict__ R, int ldr) {
  constexpr int RS = seg_row_stride(NT, NP);         // > 0: the host never emits an empty class
  constexpr int CAP = SEG_SMEM_DOUBLES / (RS > 0 ? RS : 1);
  constexpr int NPX = NP > 0 ? NP : 1, NTX = NT > 0 ? NT : 1;
  __shared__ __align__(16) double srow[SEG_SMEM_DOUBLES];
  __shared__ __align__(16) int32_t sg[CSL_GRP_STRIDE];          // the group record (host-built, CSL_G_*)
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
  const int32_t* __restrict__ grec = p.grp_tab + (long long)(bin0 / CSL_SEG_GROUP + blockIdx.y) * CSL_GRP_STRIDE;
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
  const int split = __syncthreads_or(mode[0] != mode[1]);

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
        seg_cols<NT, NP, MODE>(srow + (s0 - p0) * RS, s1 - s0, cf, pamp, pidx, E, at, ap, ltab, l, lend);
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

template <int NT, int NP>
static int launch_seg(const csl_program_t* p, int bin0, int nbins, int W, const double* coef, int ldc,
                      double* R, int ldr, cudaStream_t s) {
  dim3 grid((W + SEG_WPT * SEG_TPB - 1) / (SEG_WPT * SEG_TPB), (nbins + CSL_SEG_GROUP - 1) / CSL_SEG_GROUP);
  k_bin_segsum<NT, NP><<<grid, SEG_TPB, 0, s>>>(*p, bin0, nbins, W, coef, ldc, R, ldr);
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
static int launch_class(const csl_program_t* p, int tile0, int ntile, int W, const double* coef, int ldc,
                        double* R, int ldr, cudaStream_t s) {
  static bool attr_set = false;
  const size_t smem = (size_t)(2 * CSL_ASZ + 2 * CSL_BSZ + 2 * CSL_KC * (NT + 4 * NP)) * sizeof(double);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_bin_residual<NT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid(ntile, W / CSL_BN);
  k_bin_residual<NT, NP><<<grid, CSL_THREADS, smem, s>>>(*p, tile0, W, coef, ldc, R, ldr);
  return 0;
}

#define CSL_DISPATCH_NP(NTV)                                                                          \
  switch (np) {                                                                                        \
    case 0: return launch_class<NTV, 0>(p, tile0, ntile, W, coef, ldc, R, ldr, s);                     \
    case 1: return launch_class<NTV, 1>(p, tile0, ntile, W, coef, ldc, R, ldr, s);                     \
    case 2: return launch_class<NTV, 2>(p, tile0, ntile, W, coef, ldc, R, ldr, s);                     \
    case 4: return launch_class<NTV, 4>(p, tile0, ntile, W, coef, ldc, R, ldr, s);                     \
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: power-law class must be 0, 1, 2 or 4");  \
  }

static int dispatch_class(int nt, int np, const csl_program_t* p, int tile0, int ntile, int W, const double* coef,
                          int ldc, double* R, int ldr, cudaStream_t s) {
  switch (nt) {
    case 0: CSL_DISPATCH_NP(0)
    case 2: CSL_DISPATCH_NP(2)
    case 4: CSL_DISPATCH_NP(4)
    case 8: CSL_DISPATCH_NP(8)
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: template class must be 0, 2, 4 or 8");
  }
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
    int rc = dispatch_class(k[0], k[1], p, k[2], k[3], W, coef, ldc, R, ldr, s);
    if (rc) return rc;
  }
  for (int c = 0; c < p->nsegclass; ++c) {
    const int32_t* k = p->segcls[c];
    CSL_REQUIRE(k[0] + 2 * k[1] > 0 && p->bin_tab && p->grp_tab, "csl_bin_residual: empty segmented class");
    int rc = dispatch_seg(k[0], k[1], p, k[2], k[3], W, coef, ldc, R, ldr, s);
    if (rc) return rc;
  }
  return csl_check_launch("k_bin_residual");
}
