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

// ---- the blocked solve as independent warp-owned column slabs (no CTA barrier, no T round trip) --------------
// Forward substitution never mixes walkers: column w of Y depends on column w of R only.  A WARP therefore owns a
// slab of 16 (or 8) walkers through ALL row blocks (warp tile 64 x 16: 32 accumulators per thread), and everything
// that made k_trsm_chi2 wait disappears:
//   * the B operand of the off-diagonal update, Y_j, was written by this same warp: it comes back through a
//     warp-PRIVATE cp.async ring of [16 x 16] tiles (4 stages, ordered by __syncwarp, no CTA barrier);
//   * the ring carries on into the four chunks of R_i itself, which land exactly on the warp's private T tile (the
//     ring's four stages ARE that tile); T_i = R_i - sum_j L_ij Y_j is formed in place and is the B operand of the
//     diagonal solve at once: no global round trip, no fence, the load of R_i hidden behind the last products;
//   * the A operands -- the L_ij tiles and then the four inv(L_ii) tiles of every block, one continuous sequence --
//     reach all warps of the CTA as bulk copies (cp.async.bulk of the tile-major tables LmatT / LinvT) through a
//     ring of mbarrier-guarded stages, kept full by lane 0 of warp 0 between its own products; warps drift apart
//     by up to 4 tiles.
// Zero 8-row groups -- above the diagonal of the inv(L_ii) tiles, padding rows past n in the last block -- are
// skipped at compile time (slab_mma_chunk<LO, HI, NI>).
// Balance.  The unit of work is 8 walkers.  A launch is cut into as few rounds of (SMs x 28 units) as it needs and
// the units are spread evenly over the CTAs; a CTA of 16 consumer warps gives its first (units - 16) warps 16
// walkers and the others 8, so that at 28 units every SM sub-partition (warps w, w+4, w+8, w+12) carries 3 wide
// and 1 narrow slab = 7 units: 32 768 walkers = 4096 units = 147 CTAs x 28.  (14 wide warps per CTA left two
// sub-partitions with 4 slabs and two with 3 -- 0.461 ms against 0.438; 128-walker tiles, k_trsm_chi2_v2, left 40 of
// 148 SMs with half the work of the others -- 0.598 ms.)
// Per output element the DMMA sequence, T = R - acc and the order of the sums of squares (two partial sums per
// column: rows 0..31 and 32..63 of every block, exactly the two warp rows of k_trsm_chi2) are those of the
// kernels above: chi^2 is bitwise the same.
#define SLAB_STAGES 4
#define SLAB_MAX_WARPS 16
#define SLAB_MAX_UNITS 28
__host__ __device__ constexpr int slab_ld(int ni) { return ni == 2 ? 20 : 12; }   // row strides with conflict-free B-fragment loads
__host__ __device__ constexpr int slab_region(int ni) { return CSL_BM * slab_ld(ni) + ni * 4 * 32; }   // T tile + sums of squares

// 8-row groups LO .. HI-1 of the A tile are multiplied; the others are zero and are dropped at compile time
template <int LO, int HI, int NI>
__device__ __forceinline__ void slab_mma_chunk(double (&acc)[8][NI][2], const double* __restrict__ As,
                                               const double* __restrict__ Bs, int lane) {
  if (LO >= HI) return;
  constexpr int LD = slab_ld(NI);
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < CSL_KC / 4; ++ks) {
    double a[8], b[NI];
#pragma unroll
    for (int mi = LO; mi < HI; ++mi) a[mi] = As[(mi * 8 + g) * CSL_LDA + ks * 4 + t];
#pragma unroll
    for (int ni = 0; ni < NI; ++ni) b[ni] = Bs[(ks * 4 + t) * LD + ni * 8 + g];
#pragma unroll
    for (int mi = LO; mi < HI; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
  }
}
// runtime hi of the LAST block -> the compile-time variants (warp-uniform switch)
template <int LO, int NI>
__device__ __forceinline__ void slab_mma_chunk_hi(double (&acc)[8][NI][2], const double* __restrict__ As,
                                                  const double* __restrict__ Bs, int lane, int hi) {
  switch (hi) {
    case 1: slab_mma_chunk<LO, 1, NI>(acc, As, Bs, lane); break;
    case 2: slab_mma_chunk<LO, 2, NI>(acc, As, Bs, lane); break;
    case 3: slab_mma_chunk<LO, 3, NI>(acc, As, Bs, lane); break;
    case 4: slab_mma_chunk<LO, 4, NI>(acc, As, Bs, lane); break;
    case 5: slab_mma_chunk<LO, 5, NI>(acc, As, Bs, lane); break;
    case 6: slab_mma_chunk<LO, 6, NI>(acc, As, Bs, lane); break;
    case 7: slab_mma_chunk<LO, 7, NI>(acc, As, Bs, lane); break;
    default: slab_mma_chunk<LO, 8, NI>(acc, As, Bs, lane); break;
  }
}

// The A-tile sequence of a CTA: block i contributes its 4 i off-diagonal tiles L(i, c) and then the four tiles of
// inv(L_ii).  Lane 0 of warp 0 keeps the ring filled on the side of its own products (a 17th, producer-only warp
// would put five warps on one SM sub-partition: 5 x 32 x 112 registers no longer fit its 16 K, and 16 x 128 do).
struct slab_filler {
  int next, fi, fc;         // tile index of the next fill and its (block, chunk)
};
__device__ __forceinline__ void slab_fill_ahead(slab_filler& f, int q, const csl_program_t& p, double* Aring, uint64_t* full,
                                                uint64_t* empty) {
  const int nblk = p.n_pad / CSL_BM, nct = p.n_pad / CSL_KC;
  // tile q must be on its way before this warp waits for it; tiles up to q + STAGES - 1 are filled when their
  // stage happens to be free already (probe, no wait)
  while (f.fi < nblk && f.next <= q + SLAB_STAGES - 1) {
    const int s = f.next % SLAB_STAGES;
    if (f.next >= SLAB_STAGES) {
      const uint32_t par = ((f.next / SLAB_STAGES) - 1) & 1;
      if (f.next <= q) mbar_wait(empty + s, par);
      else if (!mbar_test(empty + s, par)) break;
    }
    const double* src = f.fc < 4 * f.fi ? p.LmatT + ((long long)f.fi * nct + f.fc) * CSL_ATILE
                                        : p.LinvT + ((long long)f.fi * 4 + (f.fc - 4 * f.fi)) * CSL_ATILE;
    mbar_arrive_expect_tx(full + s, CSL_ATILE * sizeof(double));
    bulk_g2s(Aring + s * CSL_ASZ, src, CSL_ATILE * sizeof(double), full + s);
    ++f.next;
    if (++f.fc == 4 * f.fi + 4) { f.fc = 0; ++f.fi; }
  }
}

// one warp, one slab of 8 NI walkers starting at column c0, all row blocks
template <int NI, bool FILL>
__device__ __forceinline__ void slab_consumer(const csl_program_t& p, double* R, int ldr, double* __restrict__ chi2, int c0,
                                              double* my, double* Aring, uint64_t* full, uint64_t* empty, int lane) {
  constexpr int LD = slab_ld(NI), STG = CSL_KC * LD;            // doubles of one ring stage = 16 rows of the T tile
  const int g = lane >> 2, t = lane & 3;
  const int nblk = p.n_pad / CSL_BM;
  slab_filler fill;
  fill.next = 0; fill.fi = 0; fill.fc = 0;
  // the running sums of squares of a thread are touched once per block: they live in shared memory (element e
  // of lane l at csq[32 e + l]) so that the accumulators and fragments of the products keep the registers
  double* csq = my + CSL_BM * LD + lane;
  double acc[8][NI][2];
#pragma unroll
  for (int e = 0; e < 4 * NI; ++e) csq[32 * e] = 0.0;
  // rows 16 c .. of this slab -> ring stage c % 4 (pieces of 16 bytes: 8 NI per row)
  auto issue = [&](int c) {
#pragma unroll
    for (int it = 0; it < 2 * NI; ++it) {
      const int piece = lane + 32 * it, row = piece / (4 * NI), q8 = piece % (4 * NI);
      cp_async16(my + (c % SLAB_STAGES) * STG + row * LD + 2 * q8, R + (long long)(c * CSL_KC + row) * ldr + c0 + 2 * q8);
    }
  };
  int q = 0;
  for (int i = 0; i < nblk; ++i) {
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    // rows >= n (last block only) are identity rows of L acting on zero rows of R: their 8-row groups are skipped,
    // Y stays zero there
    const int hi = min(8, (p.n - i * CSL_BM + 7) / 8);
    // acc = sum_{j<i} L_ij Y_j.  The B stream is the nup = 4 i chunks of Y rows followed by the FOUR chunks of R_i
    // itself: nup is a multiple of the ring depth, so R_i chunk c lands in stage c -- rows 16 c .. of the T tile
    const int nup = 4 * i, nall = nup + 4;
#pragma unroll
    for (int c = 0; c < SLAB_STAGES - 1; ++c) {
      issue(c);                                       // c < nall always
      cp_async_commit();
    }
    for (int c = 0; c < nup; ++c, ++q) {
      issue(c + SLAB_STAGES - 1);                     // < nall; into the stage chunk c-1 has left
      cp_async_commit();
      cp_async_wait<SLAB_STAGES - 1>();               // chunk c has landed (this lane's pieces)
      __syncwarp();                                   // ... and every other lane's
      const int s = q % SLAB_STAGES;
      if (FILL && lane == 0) slab_fill_ahead(fill, q, p, Aring, full, empty);
      mbar_wait(full + s, (q / SLAB_STAGES) & 1);
      if (hi == 8) slab_mma_chunk<0, 8, NI>(acc, Aring + s * CSL_ASZ, my + (c % SLAB_STAGES) * STG, lane);
      else slab_mma_chunk_hi<0, NI>(acc, Aring + s * CSL_ASZ, my + (c % SLAB_STAGES) * STG, lane, hi);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
    issue(nall - 1);                                  // the last chunk of R_i, into the stage the last product has left
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();
    // T_i = R_i - acc, in place in the warp's private tile (every thread rewrites exactly what it read)
#pragma unroll
    for (int mi = 0; mi < 8; ++mi)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        double2* cell = reinterpret_cast<double2*>(my + (mi * 8 + g) * LD + ni * 8 + 2 * t);
        double2 v = *cell;
        v.x = v.x - acc[mi][ni][0];
        v.y = v.y - acc[mi][ni][1];
        *cell = v;
        acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
      }
    __syncwarp();
    // Y_i = inv(L_ii) T_i : chunk c of the inverse has zero rows above 16 c
#pragma unroll
    for (int c = 0; c < 4; ++c, ++q) {
      const int s = q % SLAB_STAGES;
      if (FILL && lane == 0) slab_fill_ahead(fill, q, p, Aring, full, empty);
      mbar_wait(full + s, (q / SLAB_STAGES) & 1);
      const double* As = Aring + s * CSL_ASZ;
      const double* Bs = my + c * STG;
      if (hi == 8) {
        if (c == 0) slab_mma_chunk<0, 8, NI>(acc, As, Bs, lane);
        else if (c == 1) slab_mma_chunk<2, 8, NI>(acc, As, Bs, lane);
        else if (c == 2) slab_mma_chunk<4, 8, NI>(acc, As, Bs, lane);
        else slab_mma_chunk<6, 8, NI>(acc, As, Bs, lane);
      } else {                                     // last block: T rows >= 8 hi are zero, so are the chunks past them
        if (c == 0) slab_mma_chunk_hi<0, NI>(acc, As, Bs, lane, hi);
        else if (c == 1) slab_mma_chunk_hi<2, NI>(acc, As, Bs, lane, hi);
        else if (c == 2) slab_mma_chunk_hi<4, NI>(acc, As, Bs, lane, hi);
        else slab_mma_chunk_hi<6, NI>(acc, As, Bs, lane, hi);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int ni = 0; ni < NI; ++ni) {
        double s0 = csq[32 * (2 * NI * h + 2 * ni)], s1 = csq[32 * (2 * NI * h + 2 * ni + 1)];
#pragma unroll
        for (int mi = 4 * h; mi < 4 * h + 4; ++mi) {
          double2 y;
          y.x = acc[mi][ni][0];
          y.y = acc[mi][ni][1];
          *reinterpret_cast<double2*>(R + (long long)(i * CSL_BM + mi * 8 + g) * ldr + c0 + ni * 8 + 2 * t) = y;
          s0 = fma(y.x, y.x, s0);
          s1 = fma(y.y, y.y, s1);
        }
        csq[32 * (2 * NI * h + 2 * ni)] = s0;
        csq[32 * (2 * NI * h + 2 * ni + 1)] = s1;
      }
    // Y_i is read back by this warp's own cp.async.cg (other lanes) from the next block on.  cp.async.cg reads L2
    // past the L1: a CTA-scope fence is NOT enough (measured: about one launch in five returned a few stale rows);
    // the device-scope fence waits for L2 to hold the rows
    __threadfence();
    __syncwarp();
  }
#pragma unroll
  for (int ni = 0; ni < NI; ++ni)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      double v0 = csq[32 * (2 * ni + e)], v1 = csq[32 * (2 * NI + 2 * ni + e)];
      v0 += shfl_xor_d(v0, 4); v0 += shfl_xor_d(v0, 8); v0 += shfl_xor_d(v0, 16);
      v1 += shfl_xor_d(v1, 4); v1 += shfl_xor_d(v1, 8); v1 += shfl_xor_d(v1, 16);
      if (g == 0) chi2[c0 + ni * 8 + 2 * t + e] = v0 + v1;
    }
}

// nunit = 8-walker units of the launch, upc = units per CTA (<= 28), k = consumer warps per CTA (<= 16);
// sep = 1: warp k is a dedicated producer of the A ring (CTAs of fewer than 16 warps), 0: lane 0 of warp 0 feeds it
__global__ void __launch_bounds__(SLAB_MAX_WARPS * 32, 1)
k_trsm_chi2_slab(csl_program_t p, double* R, int ldr, double* __restrict__ chi2, int nunit, int upc, int k, int sep) {
  extern __shared__ __align__(128) double smem[];
  double* Aring = smem;                                          // [SLAB_STAGES][CSL_ASZ]
  uint64_t* full = reinterpret_cast<uint64_t*>(Aring + SLAB_STAGES * CSL_ASZ);
  uint64_t* empty = full + SLAB_STAGES;
  double* Wreg = reinterpret_cast<double*>(empty + SLAB_STAGES);   // the warps' private regions, wide ones first
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int u0 = (int)blockIdx.x * upc, uh = min(upc, nunit - u0);       // this CTA's units
  const int nwide = max(0, min(uh - k, k)), nact = min(k, uh - nwide);   // warps 0..nwide-1: 16 walkers; ..nact-1: 8
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SLAB_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, nact); }
    mbar_fence_init();
  }
  __syncthreads();
  if (sep && warp == k) {
    if (lane == 0) {
      slab_filler fill;
      fill.next = 0; fill.fi = 0; fill.fc = 0;
      const int nblk = p.n_pad / CSL_BM;
      while (fill.fi < nblk) slab_fill_ahead(fill, fill.next, p, Aring, full, empty);   // always blocking: next <= q
    }
    return;
  }
  if (warp >= nact) return;
  const int c0 = warp < nwide ? (u0 + 2 * warp) * 8 : (u0 + 2 * nwide + (warp - nwide)) * 8;
  double* my = warp < nwide ? Wreg + warp * slab_region(2) : Wreg + nwide * slab_region(2) + (warp - nwide) * slab_region(1);
  if (warp == 0 && !sep) {                                       // warp 0 also feeds the A ring
    if (nwide > 0) slab_consumer<2, true>(p, R, ldr, chi2, c0, my, Aring, full, empty, lane);
    else slab_consumer<1, true>(p, R, ldr, chi2, c0, my, Aring, full, empty, lane);
  } else if (warp < nwide) {
    slab_consumer<2, false>(p, R, ldr, chi2, c0, my, Aring, full, empty, lane);
  } else {
    slab_consumer<1, false>(p, R, ldr, chi2, c0, my, Aring, full, empty, lane);
  }
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

// chi^2 left as per-32-row partial sums in ws_part[i, :] (explicit-inverse mode; csl_combine_parts adds them): returns
// CSL_E_BADARG for a program in substitution mode, whose kernel writes chi2 directly
extern "C" int csl_chi2_parts(const csl_program_t* p, int W, double* R, int ldr, double* ws_part, void* stream) {
  CSL_REQUIRE(p && p->chi2_mode == 1 && ws_part, "csl_chi2_parts: explicit-inverse programs only");
  return csl_chi2_impl(p, W, R, ldr, ws_part, ws_part, false, stream);
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
    // super-group: as many column tiles as keep their R panels (n_pad x 128 doubles each) within 48 MB of the
    // 126 MB L2.  Measured at n_pad = 640, 256 tiles (profiles/r2_trigemm_group_sweep.json): DRAM read per launch /
    // time = 166 MB / 0.402 ms at 32 tiles, 167 / 0.386 at 64, 181 / 0.388 at 96, 248 / 0.383 at 146 (the round-1
    // choice: 96 MB of panels), 675 / 0.382 at 256 -- algorithmic 164 MB.  73 tiles keep the traffic at the
    // algorithmic figure for ~1 % of the time (DRAM is never the bound here: 0.7 TB/s at worst)
    long long panel = (long long)p->n_pad * CSL_BN * sizeof(double);
    int grp = (int)((48ll << 20) / panel);
    grp = grp < 1 ? 1 : (grp > 256 ? 256 : grp);
    {   // equal super-groups: a short last group ends the launch on a thin ladder (73, 73, 73, 37 tiles: 0.403 ms
        // against 0.386 ms for 4 x 64 at 256 tiles)
      const int nt = W / CSL_BN, ng = (nt + grp - 1) / grp;
      grp = (nt + ng - 1) / ng;
    }
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
  if (shape < 0) shape = (p->LmatT && p->LinvT) ? 4 : 2;
  if (shape == 4) {
    CSL_REQUIRE(p->LmatT && p->LinvT, "csl_chi2: the slab kernel needs the tile-major factor tables");
    // units of 8 walkers: as few rounds of (SMs x 28 units) as the launch needs, the units spread evenly over the
    // CTAs; k consumer warps: 16 when a CTA carries 16 units or more (wide + narrow slabs), else the multiple of 4
    // that holds them as wide slabs
    const int nunit = W / 8, sms = csl_sm_count();
    const int rounds = (nunit + sms * SLAB_MAX_UNITS - 1) / (sms * SLAB_MAX_UNITS);
    int upc = (nunit + sms * rounds - 1) / (sms * rounds);
    upc = upc < 1 ? 1 : (upc > SLAB_MAX_UNITS ? SLAB_MAX_UNITS : upc);
    // narrow (8-walker) slabs while a warp per unit fits: twice the warps of the wide cut hide each other's latency
    int k = upc >= SLAB_MAX_WARPS ? SLAB_MAX_WARPS : upc;      // below 16 units: one narrow warp per unit + the producer warp
    if (csl_option_value("slab_warps") > 0) {            // sweeps: all-wide CTAs of this many warps
      k = csl_option_value("slab_warps");
      k = k > SLAB_MAX_WARPS ? SLAB_MAX_WARPS : k;
      upc = 2 * k > SLAB_MAX_UNITS ? SLAB_MAX_UNITS : 2 * k;
    }
    const int sep = k < SLAB_MAX_WARPS ? 1 : 0;          // a 17th warp does not fit the register file: fold the producer
    const int grid = (nunit + upc - 1) / upc;
    const int nwide = upc - k > 0 ? (upc - k < k ? upc - k : k) : 0, nnarrow = (upc - nwide < k ? upc - nwide : k) - nwide;
    const size_t smem4 = (size_t)(SLAB_STAGES * CSL_ASZ + nwide * slab_region(2) + nnarrow * slab_region(1)) * sizeof(double) +
                         2 * SLAB_STAGES * sizeof(uint64_t);
    cudaError_t e = cudaFuncSetAttribute(k_trsm_chi2_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
    k_trsm_chi2_slab<<<grid, (k + sep) * 32, smem4, s>>>(*p, R, ldr, chi2, nunit, upc, k, sep);
    return csl_check_launch("k_trsm_chi2_slab");
  }
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
    if (shape == 3) return go(k_trsm_chi2_v2<3, true>);      // CTA-scope fences: sweeps only (see k_trsm_chi2_slab: unsafe with cp.async.cg)
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
