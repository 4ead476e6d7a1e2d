// Foreground build + window binning + residual, one kernel.
//   reference: spt_lowl.py:171-182 (cl = cmb + egfs; b_i = <cl[windowrange], w_i>), :132 (r = cal*d - b),
//              spt_lowl.py:206-207 / models/clust_poisson_egfs.py:14-17 (the foreground spectra),
//              mspec_lnl.py:50-73 (the same per cross-spectrum).
// C[bins, walkers] = Wnd[bins, ell] * Cl[ell, walkers] runs on the FP64 tensor pipe (DMMA.8x8x4).
// The Cl operand never exists in HBM: each CTA builds its [16 ell x 128 walker] chunk in shared
// memory from the walker's coefficients (registers) and the templates (L1-broadcast loads).
#include "common.cuh"

__global__ void __launch_bounds__(CSL_THREADS, 2)
k_bin_residual(csl_program_t p, int W, const double* __restrict__ coef, int ldc, double* __restrict__ R,
               int ldr) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                       // [2][CSL_ASZ]
  double* Bs = smem + 2 * CSL_ASZ;         // [2][CSL_BSZ]
  const int32_t* tile = p.tile_tab + (long long)blockIdx.x * CSL_TILE_STRIDE;
  const int32_t* spec = p.spec_tab + (long long)tile[CSL_T_SPEC] * CSL_SPEC_STRIDE;
  const int c0 = blockIdx.y * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;

  const int l_begin = tile[CSL_T_LBEGIN], l_end = tile[CSL_T_LEND];
  const long long win_off = ((long long)tile[CSL_T_WINOFF_HI] << 32) | (unsigned)tile[CSL_T_WINOFF_LO];
  const long long win_ld = tile[CSL_T_WINLD];
  const double* Wnd = p.windows + win_off;

  // builder role: this thread owns walker c0 + wl and 8 of the 16 ell rows of a chunk
  const int wl = tid & (CSL_BN - 1), grp = tid >> 7;
  const int nterm = spec[CSL_S_NTERM], nplaw = spec[CSL_S_NPLAW];
  double cf[CSL_MAX_TERMS];
  int toff[CSL_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < CSL_MAX_TERMS; ++t) {
    cf[t] = (t < nterm) ? coef[(long long)spec[CSL_S_TERM_COEF + t] * ldc + c0 + wl] : 0.0;
    toff[t] = (t < nterm) ? spec[CSL_S_TERM_OFF + t] : 0;
  }
  double pamp[CSL_MAX_PLAW], pidx[CSL_MAX_PLAW];
#pragma unroll
  for (int q = 0; q < CSL_MAX_PLAW; ++q) {
    pamp[q] = (q < nplaw) ? coef[(long long)spec[CSL_S_PLAW_AMP + q] * ldc + c0 + wl] : 0.0;
    pidx[q] = (q < nplaw) ? coef[(long long)spec[CSL_S_PLAW_IDX + q] * ldc + c0 + wl] : 0.0;
  }

  auto build_B = [&](double* B, int l0) {
#pragma unroll
    for (int r = 0; r < CSL_KC / 2; ++r) {
      const int kk = grp * (CSL_KC / 2) + r;
      const int l = l0 + kk;
      double v = 0.0;
#pragma unroll
      for (int t = 0; t < CSL_MAX_TERMS; ++t)
        if (t < nterm) v = fma(cf[t], __ldg(p.templates + toff[t] + l), v);
      for (int q = 0; q < nplaw; ++q) {
        double lb = __ldg(p.templates + spec[CSL_S_PLAW_LOGB + q] + l);
        double mu = __ldg(p.templates + spec[CSL_S_PLAW_MUL + q] + l);
        v = fma(pamp[q] * mu, exp(pidx[q] * lb), v);
      }
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
    load_A_tile_async(As, Wnd + l_begin, win_ld, tid);
    cp_async_commit();
    build_B(Bs, l_begin);
    cp_async_wait<0>();
    __syncthreads();
    for (int c = 0; c < nchunk; ++c) {
      const int cur = c & 1;
      const int l0 = l_begin + c * CSL_KC;
      if (c + 1 < nchunk) {
        load_A_tile_async(As + (cur ^ 1) * CSL_ASZ, Wnd + l0 + CSL_KC, win_ld, tid);
        cp_async_commit();
        build_B(Bs + (cur ^ 1) * CSL_BSZ, l0 + CSL_KC);
      }
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

// rows n .. n_pad of R are zero so the padded (identity) part of L contributes nothing
__global__ void k_zero_pad_rows(double* __restrict__ R, int ldr, int n, int n_pad) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)(n_pad - n) * ldr;
  if (i < tot) R[(long long)n * ldr + i] = 0.0;
}

extern "C" int csl_bin_residual(const csl_program_t* p, int W, const double* coef, int ldc, double* R, int ldr,
                                void* stream) {
  CSL_REQUIRE(p && coef && R, "csl_bin_residual: null argument");
  CSL_REQUIRE(p->resid_mode == 1 && p->ntile > 0, "csl_bin_residual: program has no binned spectra");
  CSL_REQUIRE(W > 0 && W % CSL_BN == 0 && ldr >= W && ldc >= W && (ldr % 2) == 0,
              "csl_bin_residual: W must be a positive multiple of 128 and ldr, ldc >= W");
  static bool attr_set = false;
  const size_t smem = (size_t)(2 * CSL_ASZ + 2 * CSL_BSZ) * sizeof(double);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_bin_residual, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (p->n_pad > p->n) {
    long long tot = (long long)(p->n_pad - p->n) * ldr;
    k_zero_pad_rows<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(R, ldr, p->n, p->n_pad);
  }
  dim3 grid(p->ntile, W / CSL_BN);
  k_bin_residual<<<grid, CSL_THREADS, smem, s>>>(*p, W, coef, ldc, R, ldr);
  return csl_check_launch("k_bin_residual");
}
