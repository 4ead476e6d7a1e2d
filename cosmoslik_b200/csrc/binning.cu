// Foreground build + window binning + residual, one kernel.
//   reference: spt_lowl.py:171-182 (cl = cmb + egfs; b_i = <cl[windowrange], w_i>), :132 (r = cal*d - b),
//              spt_lowl.py:206-207 / models/clust_poisson_egfs.py:14-17 (the foreground spectra),
//              mspec_lnl.py:50-73 (the same per cross-spectrum).
// C[bins, walkers] = Wnd[bins, ell] * Cl[ell, walkers] runs on the FP64 tensor pipe (DMMA.8x8x4).
// The Cl operand never exists in HBM: each CTA builds its [16 ell x 128 walker] chunk in shared
// memory from the walker's coefficients (registers) and the spectrum's ell-table, whose 16 rows for
// the chunk are staged in shared memory by cp.async together with the window tile.
//
// ell-table row (doubles): [T_0 .. T_{NT-1}, (logB_q, M_q) for q < NP]
//   cl[l] = sum_t coef_t T_t[l] + sum_q amp_q M_q[l] exp(idx_q logB_q[l])
// The kernel is instantiated per padded class <NT, NP> so the build loop is branch-free and fully
// unrolled; the host sorts the tiles by class (csl_program_t::cls).
#include "common.cuh"

template <int NT, int NP>
__global__ void __launch_bounds__(CSL_THREADS, 2)
k_bin_residual(csl_program_t p, int tile0, int W, const double* __restrict__ coef, int ldc,
               double* __restrict__ R, int ldr) {
  constexpr int LSTRIDE = NT + 2 * NP;
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
  auto build_B = [&](double* B, const double* L) {
#pragma unroll
    for (int r = 0; r < CSL_KC / 2; ++r) {
      const int kk = grp * (CSL_KC / 2) + r;
      const double* row = L + kk * LSTRIDE;     // warp-uniform address -> shared-memory broadcast
      double v = 0.0;
#pragma unroll
      for (int t = 0; t < NT; ++t) v = fma(cf[t], row[t], v);
#pragma unroll
      for (int q = 0; q < NP; ++q) v = fma(pamp[q] * row[NT + 2 * q + 1], exp(pidx[q] * row[NT + 2 * q]), v);
      B[kk * CSL_LDB + wl] = v;
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
        out.x = __dsub_rn(__dmul_rn(cal0, d), acc[mi][ni][0]);
        out.y = __dsub_rn(__dmul_rn(cal1, d), acc[mi][ni][1]);
        *reinterpret_cast<double2*>(R + (long long)(row0 + rl) * ldr + col) = out;
      }
    }
  }
}

// ---- segmented-sum binning for block-sparse windows ----------------------------------------------------
// One thread per walker, CSL_SEG_GROUP window rows per CTA row.  For each row only its non-zero support
// [lo, hi) is visited: cl[l] is built once from the (warp-uniform) ell-table row and folded into the
// bin with one extra FMA.  No shared memory, no barriers: the kernel is bound by the FP64 pipe.
template <int NT, int NP>
__global__ void __launch_bounds__(128)
k_bin_segsum(csl_program_t p, int bin0, int nbins, int W, const double* __restrict__ coef, int ldc,
             double* __restrict__ R, int ldr) {
  constexpr int LSTRIDE = NT + 2 * NP;
  const int w = blockIdx.x * 128 + threadIdx.x;
  const int b_begin = bin0 + blockIdx.y * CSL_SEG_GROUP;
  const int b_end = min(b_begin + CSL_SEG_GROUP, bin0 + nbins);
  const int32_t* first = p.bin_tab + (long long)b_begin * CSL_BIN_STRIDE;
  const int32_t* spec = p.spec_tab + (long long)first[CSL_B_SPEC] * CSL_SPEC_STRIDE;   // one spectrum per group
  const double* __restrict__ ltab = p.templates + spec[CSL_S_LTAB];
  double cf[NT > 0 ? NT : 1], pamp[NP > 0 ? NP : 1], pidx[NP > 0 ? NP : 1];
#pragma unroll
  for (int t = 0; t < NT; ++t) cf[t] = coef[(long long)spec[CSL_S_TERM_COEF + t] * ldc + w];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    pamp[q] = coef[(long long)spec[CSL_S_PLAW_AMP + q] * ldc + w];
    pidx[q] = coef[(long long)spec[CSL_S_PLAW_IDX + q] * ldc + w];
  }
  const int calrow = spec[CSL_S_CAL];
  const double cal = calrow >= 0 ? coef[(long long)calrow * ldc + w] : 1.0;

  auto cl_at = [&](const double* __restrict__ row) {
    double v = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) v = fma(cf[t], __ldg(row + t), v);
#pragma unroll
    for (int q = 0; q < NP; ++q)
      v = fma(pamp[q] * __ldg(row + NT + 2 * q + 1), exp(pidx[q] * __ldg(row + NT + 2 * q)), v);
    return v;
  };

  for (int b = b_begin; b < b_end; ++b) {
    const int32_t* rec = p.bin_tab + (long long)b * CSL_BIN_STRIDE;
    const int lo = rec[CSL_B_LO], hi = rec[CSL_B_HI], row = rec[CSL_B_ROW];
    if (row < 0) continue;   // group padding
    const long long woff = ((long long)rec[CSL_B_WINOFF_HI] << 32) | (unsigned)rec[CSL_B_WINOFF_LO];
    const double* __restrict__ wrow = p.windows + woff;
    double acc0 = 0.0, acc1 = 0.0;
    int l = lo;
    for (; l + 1 < hi; l += 2) {          // two independent elements in flight per thread
      const double v0 = cl_at(ltab + (long long)l * LSTRIDE);
      const double v1 = cl_at(ltab + (long long)(l + 1) * LSTRIDE);
      acc0 = fma(__ldg(wrow + l), v0, acc0);
      acc1 = fma(__ldg(wrow + l + 1), v1, acc1);
    }
    if (l < hi) acc0 = fma(__ldg(wrow + l), cl_at(ltab + (long long)l * LSTRIDE), acc0);
    R[(long long)row * ldr + w] = __dsub_rn(__dmul_rn(cal, p.data[row]), acc0 + acc1);
  }
}

template <int NT, int NP>
static int launch_seg(const csl_program_t* p, int bin0, int nbins, int W, const double* coef, int ldc,
                      double* R, int ldr, cudaStream_t s) {
  dim3 grid(W / 128, (nbins + CSL_SEG_GROUP - 1) / CSL_SEG_GROUP);
  k_bin_segsum<NT, NP><<<grid, 128, 0, s>>>(*p, bin0, nbins, W, coef, ldc, R, ldr);
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
  switch (nt) {
    case 0: CSL_DISPATCH_SEG_NP(0)
    case 2: CSL_DISPATCH_SEG_NP(2)
    case 4: CSL_DISPATCH_SEG_NP(4)
    case 8: CSL_DISPATCH_SEG_NP(8)
    default: return csl_fail(CSL_E_BADARG, "csl_bin_residual: template class must be 0, 2, 4 or 8");
  }
}

// rows n .. n_pad of R are zero so the padded (identity) part of L contributes nothing
__global__ void k_zero_pad_rows(double* __restrict__ R, int ldr, int n, int n_pad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)(n_pad - n) * ldr;
  if (i < tot) R[(long long)n * ldr + i] = 0.0;
}

template <int NT, int NP>
static int launch_class(const csl_program_t* p, int tile0, int ntile, int W, const double* coef, int ldc,
                        double* R, int ldr, cudaStream_t s) {
  static bool attr_set = false;
  const size_t smem = (size_t)(2 * CSL_ASZ + 2 * CSL_BSZ + 2 * CSL_KC * (NT + 2 * NP)) * sizeof(double);
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
  CSL_REQUIRE(p->resid_mode == 1 && (p->nclass > 0 || p->nsegclass > 0) && p->nclass <= CSL_MAX_CLASS &&
                  p->nsegclass <= CSL_MAX_CLASS,
              "csl_bin_residual: program has no binned spectra");
  CSL_REQUIRE(W > 0 && W % CSL_BN == 0 && ldr >= W && ldc >= W && (ldr % 2) == 0,
              "csl_bin_residual: W must be a positive multiple of 128 and ldr, ldc >= W");
  cudaStream_t s = (cudaStream_t)stream;
  if (p->n_pad > p->n) {
    long long tot = (long long)(p->n_pad - p->n) * ldr;
    k_zero_pad_rows<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(R, ldr, p->n, p->n_pad);
  }
  for (int c = 0; c < p->nclass; ++c) {
    const int32_t* k = p->cls[c];
    CSL_REQUIRE(k[0] + 2 * k[1] > 0, "csl_bin_residual: empty spectrum class");
    int rc = dispatch_class(k[0], k[1], p, k[2], k[3], W, coef, ldc, R, ldr, s);
    if (rc) return rc;
  }
  for (int c = 0; c < p->nsegclass; ++c) {
    const int32_t* k = p->segcls[c];
    CSL_REQUIRE(k[0] + 2 * k[1] > 0 && p->bin_tab, "csl_bin_residual: empty segmented class");
    int rc = dispatch_seg(k[0], k[1], p, k[2], k[3], W, coef, ldc, R, ldr, s);
    if (rc) return rc;
  }
  return csl_check_launch("k_bin_residual");
}
