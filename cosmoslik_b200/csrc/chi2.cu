// chi2[w] = || L^-1 r_w ||^2 for a tile of 128 walkers per CTA: blocked forward substitution with
// 64-row blocks.  Off-diagonal updates  T_i = R_i - sum_{j<i} L_ij Y_j  and the diagonal solves
// Y_i = inv(L_ii) T_i  (diagonal blocks inverted once on the host) all run as FP64 DMMA tile products;
// column sums of squares are reduced with warp shuffles in a fixed order (bitwise reproducible).
//   reference: spt_lowl.py:98,133 and mspec_lnl.py:31,73  (cho_factor once; cho_solve + dot per call)
// The factor streams from L2/HBM once per walker tile (cp.async, two stages); Y_j is re-read from
// global memory (L2 resident) so any n works, including factors far larger than shared memory.
#include "common.cuh"
#include <stdlib.h>

#define CHI2_SMEM_DOUBLES (2 * CSL_ASZ + 2 * CSL_BSZ + CSL_BM * CSL_LDB + 2 * CSL_BN)

__global__ void __launch_bounds__(CSL_THREADS, 1)
k_trsm_chi2(csl_program_t p, double* R, int ldr, double* __restrict__ chi2) {
  extern __shared__ __align__(128) double smem[];
  double* As = smem;
  double* Bs = As + 2 * CSL_ASZ;
  double* Ts = Bs + 2 * CSL_BSZ;            // [64][CSL_LDB]: T_i as the B operand of the diagonal solve
  double* red = Ts + CSL_BM * CSL_LDB;      // [2][128]
  const int c0 = blockIdx.x * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  const int npad = p.n_pad, nblk = npad / CSL_BM;

  double colsq[4][2];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) colsq[ni][0] = colsq[ni][1] = 0.0;

  for (int i = 0; i < nblk; ++i) {
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;

    // acc = sum_{j<i} L_ij Y_j
    gemm_stream_AB(acc, p.Lmat + (long long)i * CSL_BM * npad, npad, R + c0, ldr, i * (CSL_BM / CSL_KC), As, Bs,
                   tid, wm0, wn0, lane);

    // T_i = R_i - acc  -> shared memory, laid out as a [64 x 128] B operand
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int rl = wm0 + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int cl = wn0 + ni * 8 + 2 * t;
        double2 r = *reinterpret_cast<const double2*>(R + (long long)(i * CSL_BM + rl) * ldr + c0 + cl);
        Ts[rl * CSL_LDB + cl] = r.x - acc[mi][ni][0];
        Ts[rl * CSL_LDB + cl + 1] = r.y - acc[mi][ni][1];
        acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      }
    }
    __syncthreads();

    // Y_i = inv(L_ii) T_i
    gemm_stream_A(acc, p.Linv_diag + (long long)i * CSL_BM * CSL_BM, CSL_BM, Ts, CSL_BM / CSL_KC, As, tid, wm0,
                  wn0, lane);

#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int rl = wm0 + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        const int cl = wn0 + ni * 8 + 2 * t;
        double2 y;
        y.x = acc[mi][ni][0];
        y.y = acc[mi][ni][1];
        *reinterpret_cast<double2*>(R + (long long)(i * CSL_BM + rl) * ldr + c0 + cl) = y;
        colsq[ni][0] = fma(y.x, y.x, colsq[ni][0]);
        colsq[ni][1] = fma(y.y, y.y, colsq[ni][1]);
      }
    }
    // Y_i must be visible to the cp.async reads of the next block rows (same CTA)
    __threadfence();
    __syncthreads();
  }

  // reduce over the 8 row groups of the warp (lane bits 2..4), then over the two warp rows
#pragma unroll
  for (int ni = 0; ni < 4; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = colsq[ni][e];
      v += shfl_xor_d(v, 4);
      v += shfl_xor_d(v, 8);
      v += shfl_xor_d(v, 16);
      if (g == 0) red[(warp >> 2) * CSL_BN + wn0 + ni * 8 + 2 * t + e] = v;
    }
  __syncthreads();
  if (tid < CSL_BN) chi2[c0 + tid] = red[tid] + red[CSL_BN + tid];
}

// ---- the blocked solve, two CTAs per SM --------------------------------------------------------------
// Same arithmetic as k_trsm_chi2 (bitwise), restructured so that nothing but the two cp.async stages lives in
// shared memory: T_i = R_i - sum_{j<i} L_ij Y_j is written straight back to R_i in global memory (L2 resident) and
// the diagonal solve Y_i = inv(L_ii) T_i is four more chunks of the SAME streamed tile product, with A = the
// inverse of the diagonal block and B = the rows of T_i read back through cp.async.  That removes the 67 KB T tile
// (and its bank-conflicting stores), so two CTAs fit on an SM and a 32 768-walker launch (256 tiles) is ONE wave
// of the 296 resident slots instead of 1.73 waves of 148.
// (A variant whose producer warp fetched the freshly written rows with cp.async.bulk was faster still -- 0.58 vs
// 0.67 ms -- but about one launch in three returned stale rows for a single 32 x 32 sub-tile, with membar.gl +
// fence.proxy.async(.global) on the writer AND the reader side: profiles/r2_trsm_bulk_race.md.  Rows written in
// the same kernel are therefore read back through the generic proxy only.)
template <int NSTAGE, bool CTA_FENCE>
__global__ void __launch_bounds__(CSL_THREADS, 2)
k_trsm_chi2_v2(csl_program_t p, double* R, int ldr, double* __restrict__ chi2) {
  extern __shared__ __align__(128) double smem[];
  double* As = smem;
  double* Bs = As + NSTAGE * CSL_ASZ;
  double* red = Bs + NSTAGE * CSL_BSZ;      // [2][128]
  const int c0 = blockIdx.x * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  const int npad = p.n_pad, nblk = npad / CSL_BM;
  double colsq[4][2];
#pragma unroll
  for (int ni = 0; ni < 4; ++ni) colsq[ni][0] = colsq[ni][1] = 0.0;
  for (int i = 0; i < nblk; ++i) {
    double acc[4][4][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // acc = sum_{j<i} L_ij Y_j
    if (NSTAGE == 3)
      gemm_stream_AB3<CSL_BM, CSL_THREADS>(acc, p.Lmat + (long long)i * CSL_BM * npad, npad, R + c0, ldr,
                                           i * (CSL_BM / CSL_KC), As, Bs, tid, wm0, wn0, lane, 0xFu, -1);
    else
      gemm_stream_AB(acc, p.Lmat + (long long)i * CSL_BM * npad, npad, R + c0, ldr, i * (CSL_BM / CSL_KC), As, Bs,
                     tid, wm0, wn0, lane);
    // T_i = R_i - acc, in place in global memory (T_0 = R_0 is there already)
    if (i > 0) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) {
        const int rl = wm0 + mi * 8 + g;
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) {
          double2* ptr = reinterpret_cast<double2*>(R + (long long)(i * CSL_BM + rl) * ldr + c0 + wn0 + ni * 8 + 2 * t);
          double2 r = *ptr;
          r.x = r.x - acc[mi][ni][0];
          r.y = r.y - acc[mi][ni][1];
          *ptr = r;
          acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        }
      }
      // T_i visible to the cp.async reads below: writers and readers are threads of ONE CTA, so a CTA-scope
      // fence + the CTA barrier order them (CTA_FENCE); the device-scope fence of round 1 also invalidates L1
      // and waits for the L2 acknowledgements, ~20 times per CTA
      if (CTA_FENCE) __threadfence_block(); else __threadfence();
      __syncthreads();
    }
    // Y_i = inv(L_ii) T_i : four more chunks of the same stream
    if (NSTAGE == 3)
      gemm_stream_AB3<CSL_BM, CSL_THREADS>(acc, p.Linv_diag + (long long)i * CSL_BM * CSL_BM, CSL_BM,
                                           R + (long long)i * CSL_BM * ldr + c0, ldr, CSL_BM / CSL_KC, As, Bs, tid, wm0,
                                           wn0, lane, 0xFu, -1);
    else
      gemm_stream_AB(acc, p.Linv_diag + (long long)i * CSL_BM * CSL_BM, CSL_BM, R + (long long)i * CSL_BM * ldr + c0, ldr,
                     CSL_BM / CSL_KC, As, Bs, tid, wm0, wn0, lane);
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      const int rl = wm0 + mi * 8 + g;
#pragma unroll
      for (int ni = 0; ni < 4; ++ni) {
        double2 y;
        y.x = acc[mi][ni][0];
        y.y = acc[mi][ni][1];
        *reinterpret_cast<double2*>(R + (long long)(i * CSL_BM + rl) * ldr + c0 + wn0 + ni * 8 + 2 * t) = y;
        colsq[ni][0] = fma(y.x, y.x, colsq[ni][0]);
        colsq[ni][1] = fma(y.y, y.y, colsq[ni][1]);
      }
    }
    if (CTA_FENCE) __threadfence_block(); else __threadfence();
    __syncthreads();
  }
#pragma unroll
  for (int ni = 0; ni < 4; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = colsq[ni][e];
      v += shfl_xor_d(v, 4);
      v += shfl_xor_d(v, 8);
      v += shfl_xor_d(v, 16);
      if (g == 0) red[(warp >> 2) * CSL_BN + wn0 + ni * 8 + 2 * t + e] = v;
    }
  __syncthreads();
  if (tid < CSL_BN) chi2[c0 + tid] = red[tid] + red[CSL_BN + tid];
}

// ---- explicit-inverse mode ---------------------------------------------------------------------------
// Y = M R with M = L^-1 (lower triangular, inverted once on the host in extended precision and verified
// against the substitution result before it is used).  Every row block of Y is an independent tile
// product over the columns up to its diagonal, so the grid is (column tiles x row blocks) with no sequential
// dependency and the same flop count as the blocked solve (n^2 + 64 n).  Each warp writes the column sums
// of squares of its 32 rows to part[row / 32, :]; the consumer (k_sum_partials, or the fused accept kernel)
// adds the n_pad / 32 partials in a fixed order.
//   ROWS = 64: 8 warps, 2 CTAs per SM -- the bulk shape.
//   ROWS = 32: 4 warps, 4 CTAs per SM, twice as many (shorter) CTAs -- for launches too small to fill the
//              GPU with 64-row tiles (a few thousand walkers).  Both shapes run the same DMMA sequence per
//              32 x 32 warp tile, so chi^2 is bitwise the same whichever is chosen.
// pipeline depth: 2 stages for the bulk shape (two resident CTAs hide each other's load latency; measured
// 0.406 vs 0.412 ms with 3), 3 for the small-launch shape (0.066 vs 0.071 ms at 4096 walkers)
#define CHI2_STAGES(ROWS) ((ROWS) == 32 ? 3 : 2)
template <int ROWS>
__global__ void __launch_bounds__(4 * ROWS, 128 / ROWS)
k_trigemm_chi2(csl_program_t p, const double* __restrict__ R, int ldr, double* __restrict__ part, int grp) {
  constexpr int THREADS = 4 * ROWS, ASZ = ROWS * CSL_LDA;
  extern __shared__ __align__(128) double smem[];
  double* As = smem;
  double* Bs = As + CHI2_STAGES(ROWS) * ASZ;
  const int npad = p.n_pad, nblk = npad / ROWS;
  // CTA order: row block i costs i+1 units, and the hardware hands CTAs out in index order, so the order
  // IS the schedule.  Column tiles are taken in super-groups of `grp` tiles whose R panels fit in L2
  // together; within a super-group all the longest row blocks go first (longest-processing-time order),
  // so the launch ends on short CTAs.  (Tile-major order ended on a full 10,9,..,1 ladder: 13 % idle.)
  const int ntile = gridDim.x / nblk;
  const int per = grp * nblk;
  const int sg = blockIdx.x / per, rem = blockIdx.x - sg * per;
  const int gcur = min(grp, ntile - sg * grp);
  const int i = nblk - 1 - rem / gcur, ct = sg * grp + rem % gcur;
  const int c0 = ct * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  // rows >= n of M are identity rows acting on zero rows of R: skip their 8-row groups altogether
  unsigned mmask = 0;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
    if (i * ROWS + wm0 + mi * 8 < p.n) mmask |= 1u << mi;
  const int kend = min((i + 1) * ROWS, (p.n + CSL_KC - 1) / CSL_KC * CSL_KC);
  if (CHI2_STAGES(ROWS) == 3)
    gemm_stream_AB3<ROWS, THREADS>(acc, p.Minv + (long long)i * ROWS * npad, npad, R + c0, ldr, kend / CSL_KC, As, Bs,
                                   tid, wm0, wn0, lane, mmask, i * ROWS);
  else
    gemm_stream_AB<ROWS, THREADS>(acc, p.Minv + (long long)i * ROWS * npad, npad, R + c0, ldr, kend / CSL_KC, As, Bs,
                                  tid, wm0, wn0, lane, mmask, i * ROWS);
  double* prow = part + (long long)((i * ROWS + wm0) / 32) * ldr + c0 + wn0;
#pragma unroll
  for (int ni = 0; ni < 4; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = 0.0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) v = fma(acc[mi][ni][e], acc[mi][ni][e], v);
      v += shfl_xor_d(v, 4);
      v += shfl_xor_d(v, 8);
      v += shfl_xor_d(v, 16);
      if (g == 0) prow[ni * 8 + 2 * t + e] = v;
    }
}

// ---- the same tile product with Blackwell tile movement ------------------------------------------------
// One PRODUCER warp feeds a ring of STAGES stages with bulk asynchronous copies (cp.async.bulk, the TMA unit):
// per k-chunk ONE copy of the tile-major M tile (ROWS x 20 doubles, contiguous in MinvT) and 16 copies of
// 1 KB R rows into the padded B stage, all completing on the stage's `full` mbarrier (expect_tx).  The
// CONSUMER warps wait on `full`, run the DMMA chunk, and arrive on `empty`; they never compute a global address,
// issue a load or meet a CTA-wide barrier inside the loop, and a warp may run up to STAGES chunks ahead of its
// neighbours.  Per 32 x 32 warp tile the DMMA sequence is the one of k_trigemm_chi2, so chi^2 is bitwise the same.
template <int ROWS, int STAGES>
__global__ void __launch_bounds__(4 * ROWS + 32, ROWS == 64 ? 2 : 3)
k_trigemm_chi2_bulk(csl_program_t p, const double* __restrict__ R, int ldr, double* __restrict__ part, int grp) {
  constexpr int NCW = ROWS / 8;                          // consumer warps (each 32 x 32 of the ROWS x 128 tile)
  constexpr int ASZ = ROWS * CSL_LDA, STAGE = ASZ + CSL_BSZ;
  constexpr uint32_t TX_BYTES = (uint32_t)(ASZ + CSL_KC * CSL_BN) * sizeof(double);
  extern __shared__ __align__(128) double smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  const int npad = p.n_pad, nblk = npad / ROWS;
  const int ntile = gridDim.x / nblk;
  const int per = grp * nblk;
  const int sg = blockIdx.x / per, rem = blockIdx.x - sg * per;
  const int gcur = min(grp, ntile - sg * grp);
  const int i = nblk - 1 - rem / gcur, ct = sg * grp + rem % gcur;
  const int c0 = ct * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kend = min((i + 1) * ROWS, (p.n + CSL_KC - 1) / CSL_KC * CSL_KC);
  const int nchunk = kend / CSL_KC;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NCW); }
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == NCW) {
    // ---------------- producer
    const int nct = npad / CSL_KC;
    const double* __restrict__ At = p.MinvT + ((long long)((i * ROWS) / CSL_BM) * nct) * CSL_ATILE + ((i * ROWS) % CSL_BM) * CSL_LDA;
    const double* __restrict__ Bg = R + c0;
    for (int c = 0; c < nchunk; ++c) {
      const int s = c % STAGES, use = c / STAGES;
      if (use > 0) mbar_wait(empty + s, (use - 1) & 1);          // the consumers have left this stage
      double* As = smem + s * STAGE;
      double* Bs = As + ASZ;
      if (lane == 0) mbar_arrive_expect_tx(full + s, TX_BYTES);
      __syncwarp();
      if (lane < CSL_KC)
        bulk_g2s(Bs + lane * CSL_LDB, Bg + (long long)(c * CSL_KC + lane) * ldr, CSL_BN * sizeof(double), full + s);
      else if (lane == CSL_KC)
        bulk_g2s(As, At + (long long)c * CSL_ATILE, ASZ * sizeof(double), full + s);
    }
    return;
  }
  // ---------------- consumers
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  // rows >= n of M are identity rows acting on zero rows of R: skip their 8-row groups altogether
  unsigned mmask = 0;
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
    if (i * ROWS + wm0 + mi * 8 < p.n) mmask |= 1u << mi;
  for (int c = 0; c < nchunk; ++c) {
    const int s = c % STAGES;
    unsigned m = mmask;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
      if (i * ROWS + wm0 + mi * 8 + 7 < c * CSL_KC) m &= ~(1u << mi);      // 8-row group entirely above the diagonal
    mbar_wait(full + s, (c / STAGES) & 1);
    const double* As = smem + s * STAGE;
    if (m) warp_mma_chunk(acc, As, As + ASZ, wm0, wn0, lane, m);
    __syncwarp();
    if (lane == 0) mbar_arrive(empty + s);
  }
  double* prow = part + (long long)((i * ROWS + wm0) / 32) * ldr + c0 + wn0;
#pragma unroll
  for (int ni = 0; ni < 4; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v = 0.0;
#pragma unroll
      for (int mi = 0; mi < 4; ++mi) v = fma(acc[mi][ni][e], acc[mi][ni][e], v);
      v += shfl_xor_d(v, 4);
      v += shfl_xor_d(v, 8);
      v += shfl_xor_d(v, 16);
      if (g == 0) prow[ni * 8 + 2 * t + e] = v;
    }
}

template <int ROWS, int STAGES>
static int launch_trigemm_bulk(const csl_program_t* p, int nctas, const double* R, int ldr, double* part, int grp,
                               cudaStream_t s) {
  const size_t smem = (size_t)STAGES * (ROWS * CSL_LDA + CSL_BSZ) * sizeof(double) + 2 * STAGES * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(k_trigemm_chi2_bulk<ROWS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  k_trigemm_chi2_bulk<ROWS, STAGES><<<nctas, 4 * ROWS + 32, smem, s>>>(*p, R, ldr, part, grp);
  return 0;
}

__global__ void k_sum_partials(int nblk, int W, int ldr, const double* __restrict__ part, double* __restrict__ chi2) {
  int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double s = 0.0;
  for (int i = 0; i < nblk; ++i) s += part[(long long)i * ldr + w];
  chi2[w] = s;
}

extern "C" int csl_chi2(const csl_program_t* p, int W, double* R, int ldr, double* chi2, double* ws_part,
                        void* stream) {
  return csl_chi2_impl(p, W, R, ldr, chi2, ws_part, true, stream);
}

// sum_partials = false (explicit-inverse mode only): leave the per-row-block column sums in ws_part[i, :]
// for a consumer that adds them itself (the fused accept kernel of csl_stretch_run)
int csl_chi2_impl(const csl_program_t* p, int W, double* R, int ldr, double* chi2, double* ws_part, bool sum_partials,
                  void* stream) {
  CSL_REQUIRE(p && R && chi2, "csl_chi2: null argument");
  CSL_REQUIRE(p->n_pad > 0 && p->n_pad % CSL_BM == 0, "csl_chi2: program has no factor");
  CSL_REQUIRE(W > 0 && W % CSL_BN == 0 && ldr >= W && (ldr % 2) == 0, "csl_chi2: W must be a positive multiple of 128");
  cudaStream_t s = (cudaStream_t)stream;
  if (p->chi2_mode == 1) {
    CSL_REQUIRE(p->Minv && ws_part, "csl_chi2: explicit-inverse mode needs Minv and the partial-sum workspace");
    // super-group: as many column tiles as keep their R panels (n_pad x 128 doubles each) within ~96 MB
    // (measured at n_pad = 640, 256 tiles: tile-major 0.445 ms, 16: 0.45, 64: 0.409, 256: 0.408)
    long long panel = (long long)p->n_pad * CSL_BN * sizeof(double);
    int grp = (int)((96ll << 20) / panel);
    grp = grp < 1 ? 1 : (grp > 256 ? 256 : grp);
    if (csl_option_value("trigemm_group") > 0) grp = csl_option_value("trigemm_group");
    // tile shape: 64-row CTAs (2 per SM) unless the launch would leave SMs idle -- then 32-row CTAs: twice as
    // many, half as long, so the last wave is shorter.  Work of a launch in 64-row-block units: every column
    // tile carries nblk (nblk + 1) / 2 of them.
    const int nblk64 = p->n_pad / CSL_BM, ntile = W / CSL_BN;
    const int sms = csl_sm_count();
    int rows = (long long)ntile * nblk64 < 6ll * sms ? 32 : 64;
    if (csl_option_value("trigemm_rows") == 32 || csl_option_value("trigemm_rows") == 64) rows = csl_option_value("trigemm_rows");
    const int pipe = csl_option_value("pipe");
    const bool bulk = p->MinvT != nullptr && pipe != 0;
    int rc = 0;
    if (bulk) {
      const int st = csl_option_value("trigemm_stages");
      if (rows == 64) rc = st == 3 ? launch_trigemm_bulk<64, 3>(p, ntile * nblk64, R, ldr, ws_part, grp, s)
                                   : launch_trigemm_bulk<64, 4>(p, ntile * nblk64, R, ldr, ws_part, grp, s);
      else rc = st == 4 ? launch_trigemm_bulk<32, 4>(p, ntile * nblk64 * 2, R, ldr, ws_part, grp, s)     // 2 CTAs / SM
                        : launch_trigemm_bulk<32, 3>(p, ntile * nblk64 * 2, R, ldr, ws_part, grp, s);    // 3 CTAs / SM
      if (rc) return rc;
    } else if (rows == 64) {
      const size_t smem1 = (size_t)CHI2_STAGES(64) * (CSL_ASZ + CSL_BSZ) * sizeof(double);
      cudaError_t e = cudaFuncSetAttribute(k_trigemm_chi2<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
      if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
      k_trigemm_chi2<64><<<ntile * nblk64, 256, smem1, s>>>(*p, R, ldr, ws_part, grp);
    } else {
      const size_t smem1 = (size_t)CHI2_STAGES(32) * (32 * CSL_LDA + CSL_BSZ) * sizeof(double);
      cudaError_t e = cudaFuncSetAttribute(k_trigemm_chi2<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
      if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
      k_trigemm_chi2<32><<<ntile * nblk64 * 2, 128, smem1, s>>>(*p, R, ldr, ws_part, grp);
    }
    if (sum_partials) k_sum_partials<<<(W + 255) / 256, 256, 0, s>>>(2 * nblk64, W, ldr, ws_part, chi2);
    return csl_check_launch("k_trigemm_chi2");
  }
  CSL_REQUIRE(p->Lmat && p->Linv_diag, "csl_chi2: program has no factor");
  // shape: 0 = the round-1 kernel (one CTA per SM, T tile in shared memory); 1 / 2 = two CTAs per SM with a 2- / 3-stage
  // cp.async ring and device-scope fences; 3 (default) = 3 stages and CTA-scope fences.  Measured at n = 613 on
  // 4096 / 32 768 / 65 536 walkers: shape 0 0.332 / 0.668 / 1.331 ms, shape 2 0.338 / 0.600 / 1.164, shape 3
  // 0.323 / 0.596 / 1.151 (58 % / 60 % of the FP64 peak at the two large sizes); all bitwise equal
  int shape = csl_option_value("trsm_shape");
  if (shape < 0) shape = 3;
  if (shape != 0) {
    const int nst = shape == 1 ? 2 : 3;
    const size_t smem2 = (size_t)(nst * (CSL_ASZ + CSL_BSZ) + 2 * CSL_BN) * sizeof(double);
    auto go = [&](auto kern) -> int {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
      kern<<<W / CSL_BN, CSL_THREADS, smem2, s>>>(*p, R, ldr, chi2);
      return csl_check_launch("k_trsm_chi2_v2");
    };
    if (shape == 1) return go(k_trsm_chi2_v2<2, false>);
    if (shape == 3) return go(k_trsm_chi2_v2<3, true>);
    return go(k_trsm_chi2_v2<3, false>);
  }
  const size_t smem = (size_t)CHI2_SMEM_DOUBLES * sizeof(double);
  {
    cudaError_t e = cudaFuncSetAttribute(k_trsm_chi2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  k_trsm_chi2<<<W / CSL_BN, CSL_THREADS, smem, s>>>(*p, R, ldr, chi2);
  return csl_check_launch("k_trsm_chi2");
}

// ---- plain FP64 DMMA GEMM on the same tile core (peak / unit tests) ----------------------------------
__global__ void __launch_bounds__(CSL_THREADS, 2)
k_dgemm(int M, int N, int K, const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C) {
  extern __shared__ __align__(128) double smem[];
  double* As = smem;
  double* Bs = As + 2 * CSL_ASZ;
  const int m0 = blockIdx.x * CSL_BM, c0 = blockIdx.y * CSL_BN;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 32;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
  gemm_stream_AB(acc, A + (long long)m0 * K, K, B + c0, N, K / CSL_KC, As, Bs, tid, wm0, wn0, lane);
#pragma unroll
  for (int mi = 0; mi < 4; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      double2 y;
      y.x = acc[mi][ni][0];
      y.y = acc[mi][ni][1];
      *reinterpret_cast<double2*>(C + (long long)(m0 + wm0 + mi * 8 + g) * N + c0 + wn0 + ni * 8 + 2 * t) = y;
    }
}

extern "C" int csl_dgemm_dmma(int M, int N, int K, const double* A, const double* B, double* C, void* stream) {
  CSL_REQUIRE(A && B && C, "csl_dgemm_dmma: null argument");
  CSL_REQUIRE(M > 0 && N > 0 && K > 0 && M % CSL_BM == 0 && N % CSL_BN == 0 && K % CSL_KC == 0,
              "csl_dgemm_dmma: M%64, N%128, K%16 must be 0");
  const size_t smem = (size_t)(2 * CSL_ASZ + 2 * CSL_BSZ) * sizeof(double);
  {
    cudaError_t e = cudaFuncSetAttribute(k_dgemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  dim3 grid(M / CSL_BM, N / CSL_BN);
  k_dgemm<<<grid, CSL_THREADS, smem, (cudaStream_t)stream>>>(M, N, K, A, B, C);
  return csl_check_launch("k_dgemm");
}
