// Shared device helpers for the cosmoslik_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/cosmoslik_b200.h"

#define CSL_THREADS 256
#define CSL_LDA (CSL_KC + 4)   // 20 doubles: A-fragment loads hit 16 distinct bank pairs
#define CSL_LDB (CSL_BN + 4)   // 132 doubles: same for B fragments

extern thread_local char csl_err_buf[256];
int csl_fail(int code, const char* msg);
int csl_check_launch(const char* what);

// internal (C++ linkage, not in the C ABI): fused stages of csl_stretch_run and the chi^2 partial sums
struct csl_peer_args {            // exchange ordering of a sharded ensemble folded into a fused stage (samplers.cu)
  const uint64_t* flag_ptrs;      // device array [world] of every rank's flag array
  uint32_t* sync;                 // device uint32[2]: sticky status word, ticket counter
  uint64_t epoch;                 // wait: epoch every peer must have reached; signal: epoch announced
  int rank, world;
};
long long csl_barrier_timeout_cycles();
int csl_peer_signal(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, void* stream);
int csl_fused_propose_prepare(const csl_program_t* p, int Ns, int Wp, int Nc, const double* s, const double* comp,
                              double a, uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0, double* q,
                              double* zz, double* coef, int ldc, double* prior, const csl_peer_args* wait,
                              void* stream);
int csl_fused_combine_accept(const csl_program_t* p, int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q,
                             double* lnl_q, const double* zz, uint64_t seed, uint32_t step, uint32_t half,
                             uint32_t walker0, uint8_t* accepted, uint64_t* naccepted, const uint64_t* peer_ptrs,
                             int npeers, uint64_t multicast_ptr, int64_t dst_row0, const double* coef, int ldc,
                             const double* prior, const double* chi2, const double* part, int nblk, int ldp,
                             const csl_peer_args* signal, double* weight, double* rows, int32_t* nrows, int cap,
                             int last_step, void* stream);
int csl_chi2_impl(const csl_program_t* p, int W, double* R, int ldr, double* chi2, double* ws_part, bool sum_partials,
                  void* stream);
// stage variants with the padded walker count (columns written) separate from the workspace stride (prepare.cu)
int csl_eval_prepare_ld(const csl_program_t* p, int W, int Wp, const double* x, double* coef, int ldc, double* prior,
                        void* stream);
int csl_direct_residual_ld(const csl_program_t* p, int W, int Wp, const double* coef, int ldc, double* R, int ldr,
                           void* stream);
// device / environment facts read once per process (api.cu)
int csl_sm_count();
int csl_env_int(const char* name, int fallback);
int csl_option_value(const char* name);     // tuning option (csl_set_option / environment), -1 = unset
int csl_small_stage_bytes(const csl_program_t* p);   // shared-memory bytes of the staged bytecode / prior tables, 0 = not staged

#define CSL_REQUIRE(cond, msg) do { if (!(cond)) return csl_fail(CSL_E_BADARG, msg); } while (0)

// D(8x8) += A(8x4) * B(4x8) in FP64 on the tensor pipe (SASS: DMMA.8x8x4).
//   a: A[row = lane>>2][k = lane&3]    b: B[k = lane&3][col = lane>>2]
//   c0,c1: C[row = lane>>2][col = 2*(lane&3) + {0,1}]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" :: "r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" :: "n"(N)); }

// ---- Blackwell tile movement: bulk asynchronous copies (the TMA unit; SASS UBLKCP) completing on an mbarrier
// (SASS SYNCS).  One elected lane issues a copy of a whole contiguous tile -- the A operands are stored TILE-MAJOR
// by the host for exactly this: a [64 x 16] tile with its 4-double row padding is 10 240 contiguous bytes -- the
// consumers wait on the stage's "full" barrier, and release the stage by arriving on its "empty" barrier.  No
// thread of a consumer warp computes an address, issues a load or takes part in a CTA-wide barrier in the main loop.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// non-blocking probe of a barrier phase
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// global -> shared bulk copy of `bytes` (multiple of 16; both addresses 16-byte aligned), completing on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#define CSL_ATILE (CSL_BM * CSL_LDA)     // doubles of one tile-major A tile: 64 rows x (16 + 4 padding)

// One k-chunk (CSL_KC columns) of the CTA tile product on a warp tile of 32 x 32:
// acc[mi][ni][2] += A[wm0 + 8 mi .., :] * B[:, wn0 + 8 ni ..]
// As: [CSL_BM][CSL_LDA], Bs: [CSL_KC][CSL_LDB].  Bit mi of MASK = 0 drops an all-zero 8-row group at
// COMPILE time: a predicated-off DMMA still occupies the FP64 pipe (measured), a removed one does not.
template <unsigned MASK>
__device__ __forceinline__ void warp_mma_chunk_m(double (&acc)[4][4][2], const double* __restrict__ As,
                                                 const double* __restrict__ Bs, int wm0, int wn0, int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int ks = 0; ks < CSL_KC / 4; ++ks) {
    double a[4], b[4];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
      if (MASK & (1u << mi)) a[mi] = As[(wm0 + mi * 8 + g) * CSL_LDA + ks * 4 + t];
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) b[ni] = Bs[(ks * 4 + t) * CSL_LDB + wn0 + ni * 8 + g];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) {
      if (MASK & (1u << mi)) {
#pragma unroll
        for (int ni = 0; ni < 4; ++ni) dmma884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
      }
    }
  }
}

// runtime mask -> one of the 16 compile-time variants (warp-uniform switch, a real branch)
__device__ __forceinline__ void warp_mma_chunk(double (&acc)[4][4][2], const double* __restrict__ As,
                                               const double* __restrict__ Bs, int wm0, int wn0, int lane,
                                               unsigned mmask = 0xFu) {
  switch (mmask & 0xFu) {
    case 0xF: warp_mma_chunk_m<0xF>(acc, As, Bs, wm0, wn0, lane); break;
    case 0xE: warp_mma_chunk_m<0xE>(acc, As, Bs, wm0, wn0, lane); break;
    case 0xC: warp_mma_chunk_m<0xC>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x8: warp_mma_chunk_m<0x8>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x7: warp_mma_chunk_m<0x7>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x3: warp_mma_chunk_m<0x3>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x1: warp_mma_chunk_m<0x1>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x6: warp_mma_chunk_m<0x6>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x4: warp_mma_chunk_m<0x4>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x2: warp_mma_chunk_m<0x2>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x0: break;
    case 0x5: warp_mma_chunk_m<0x5>(acc, As, Bs, wm0, wn0, lane); break;
    case 0x9: warp_mma_chunk_m<0x9>(acc, As, Bs, wm0, wn0, lane); break;
    case 0xA: warp_mma_chunk_m<0xA>(acc, As, Bs, wm0, wn0, lane); break;
    case 0xB: warp_mma_chunk_m<0xB>(acc, As, Bs, wm0, wn0, lane); break;
    default: warp_mma_chunk_m<0xD>(acc, As, Bs, wm0, wn0, lane); break;
  }
}

// cp.async a [ROWS x CSL_KC] tile of a row-major matrix (row stride ld doubles) into As, by THREADS threads.
template <int ROWS = CSL_BM, int THREADS = CSL_THREADS>
__device__ __forceinline__ void load_A_tile_async(double* As, const double* __restrict__ A, long long ld,
                                                  int tid) {
  // ROWS rows x 8 chunks of 16 bytes, 2 per thread for (64, 256) and (32, 128)
  static_assert((ROWS * CSL_KC / 2) % THREADS == 0, "A tile must divide over the threads");
#pragma unroll
  for (int it = 0; it < (ROWS * CSL_KC / 2) / THREADS; ++it) {
    int c = tid + it * THREADS;
    int r = c >> 3, q = c & 7;
    cp_async16(As + r * CSL_LDA + q * 2, A + (long long)r * ld + q * 2);
  }
}

// cp.async a [CSL_KC x CSL_BN] tile (row stride ld doubles) into Bs.
template <int THREADS = CSL_THREADS>
__device__ __forceinline__ void load_B_tile_async(double* Bs, const double* __restrict__ B, long long ld,
                                                  int tid) {
  // 16 rows x 64 chunks = 1024 chunks, 4 (8) per thread
#pragma unroll
  for (int it = 0; it < (CSL_KC * CSL_BN / 2) / THREADS; ++it) {
    int c = tid + it * THREADS;
    int r = c >> 6, q = c & 63;
    cp_async16(Bs + r * CSL_LDB + q * 2, B + (long long)r * ld + q * 2);
  }
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

#define CSL_ASZ (CSL_BM * CSL_LDA)   // doubles per A stage
#define CSL_BSZ (CSL_KC * CSL_LDB)   // doubles per B stage

// acc += A[ROWS, 16*nchunk] * B[16*nchunk, 128], both operands streamed from global memory through a
// two-stage cp.async pipeline (A stage = ROWS * CSL_LDA doubles).  All threads of the CTA (THREADS = 4 ROWS:
// one warp per 32 x 32 piece of the tile) call this; on return the stages are free.
template <int ROWS = CSL_BM, int THREADS = CSL_THREADS>
__device__ __forceinline__ void gemm_stream_AB(double (&acc)[4][4][2], const double* __restrict__ A,
                                               long long lda, const double* __restrict__ B, long long ldb,
                                               int nchunk, double* As, double* Bs, int tid, int wm0, int wn0,
                                               int lane, unsigned mmask = 0xFu, int tri_row0 = -1) {
  // tri_row0 >= 0: A is lower triangular and this tile's first row is global row tri_row0 (columns start
  // at 0): 8-row groups that lie entirely above the diagonal of a chunk are skipped.
  constexpr int ASZ = ROWS * CSL_LDA;
  if (nchunk <= 0) return;
  load_A_tile_async<ROWS, THREADS>(As, A, lda, tid);
  load_B_tile_async<THREADS>(Bs, B, ldb, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  for (int c = 0; c < nchunk; ++c) {
    const int cur = c & 1;
    if (c + 1 < nchunk) {
      load_A_tile_async<ROWS, THREADS>(As + (cur ^ 1) * ASZ, A + (long long)(c + 1) * CSL_KC, lda, tid);
      load_B_tile_async<THREADS>(Bs + (cur ^ 1) * CSL_BSZ, B + (long long)(c + 1) * CSL_KC * ldb, ldb, tid);
      cp_async_commit();
    }
    unsigned m = mmask;
    if (tri_row0 >= 0) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
        if (tri_row0 + wm0 + mi * 8 + 7 < c * CSL_KC) m &= ~(1u << mi);
    }
    if (m) warp_mma_chunk(acc, As + cur * ASZ, Bs + cur * CSL_BSZ, wm0, wn0, lane, m);
    cp_async_wait<0>();
    __syncthreads();
  }
}

// The same product through a THREE-stage pipeline: the load of chunk c+2 is issued before chunk c is
// multiplied, so a chunk has two multiply-times to arrive (one is not enough when a CTA runs alone on its SM,
// as in launches of a few thousand walkers), still with one CTA barrier per chunk.
// As / Bs hold 3 stages.
template <int ROWS = CSL_BM, int THREADS = CSL_THREADS>
__device__ __forceinline__ void gemm_stream_AB3(double (&acc)[4][4][2], const double* __restrict__ A,
                                                long long lda, const double* __restrict__ B, long long ldb,
                                                int nchunk, double* As, double* Bs, int tid, int wm0, int wn0,
                                                int lane, unsigned mmask, int tri_row0) {
  constexpr int ASZ = ROWS * CSL_LDA;
  if (nchunk <= 0) return;
  auto load = [&](int c) {
    if (c < nchunk) {
      const int st = c % 3;
      load_A_tile_async<ROWS, THREADS>(As + st * ASZ, A + (long long)c * CSL_KC, lda, tid);
      load_B_tile_async<THREADS>(Bs + st * CSL_BSZ, B + (long long)c * CSL_KC * ldb, ldb, tid);
    }
    cp_async_commit();                       // always a group, possibly empty: wait_group counts groups
  };
  load(0);
  load(1);
  for (int c = 0; c < nchunk; ++c) {
    cp_async_wait<1>();                      // chunk c has landed (all but the newest group)
    __syncthreads();                         // ... for every thread; and chunk c-1 is consumed by all
    load(c + 2);                             // into the stage chunk c-1 occupied
    unsigned m = mmask;
    if (tri_row0 >= 0) {
#pragma unroll
      for (int mi = 0; mi < 4; ++mi)
        if (tri_row0 + wm0 + mi * 8 + 7 < c * CSL_KC) m &= ~(1u << mi);
    }
    const int st = c % 3;
    if (m) warp_mma_chunk(acc, As + st * ASZ, Bs + st * CSL_BSZ, wm0, wn0, lane, m);
  }
  cp_async_wait<0>();
}

// Same with the B operand already resident in shared memory as [16*nchunk][CSL_LDB].
__device__ __forceinline__ void gemm_stream_A(double (&acc)[4][4][2], const double* __restrict__ A,
                                              long long lda, const double* Bsm, int nchunk, double* As,
                                              int tid, int wm0, int wn0, int lane) {
  load_A_tile_async(As, A, lda, tid);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  for (int c = 0; c < nchunk; ++c) {
    const int cur = c & 1;
    if (c + 1 < nchunk) {
      load_A_tile_async(As + (cur ^ 1) * CSL_ASZ, A + (long long)(c + 1) * CSL_KC, lda, tid);
      cp_async_commit();
    }
    warp_mma_chunk(acc, As + cur * CSL_ASZ, Bsm + c * CSL_BSZ, wm0, wn0, lane);
    cp_async_wait<0>();
    __syncthreads();
  }
}
