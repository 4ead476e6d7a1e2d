// C-ABI glue: error reporting, device info and the composed posterior evaluation.
#include <string.h>
#include "eval_device.cuh"
#include <stdlib.h>

thread_local char csl_err_buf[256] = "";

int csl_fail(int code, const char* msg) {
  snprintf(csl_err_buf, sizeof(csl_err_buf), "%s", msg ? msg : "unknown error");
  return code;
}

int csl_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(csl_err_buf, sizeof(csl_err_buf), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

extern "C" int csl_abi_version(void) { return CSL_ABI_VERSION; }
extern "C" const char* csl_last_error(void) { return csl_err_buf; }
extern "C" int csl_sizeof_program(void) { return (int)sizeof(csl_program_t); }
extern "C" int csl_sizeof_ensemble(void) { return (int)sizeof(csl_ensemble_t); }

extern "C" int csl_device_info(int device, int32_t* out) {
  CSL_REQUIRE(out, "csl_device_info: null output");
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return csl_fail(CSL_E_NODEVICE, cudaGetErrorString(e));
  out[0] = prop.multiProcessorCount;
  out[1] = prop.major;
  out[2] = prop.minor;
  out[3] = (int32_t)prop.sharedMemPerBlockOptin;
  return 0;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// SM count of the CURRENT device, cached per device (launch-shape decisions)
int csl_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!cached[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

// integer environment override; callers cache the result in a function-local static
int csl_env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : fallback;
}

// bytecode + prior tables are interpreted from shared memory when they fit the staging area (eval_device.cuh);
// CSL_NO_STAGE_TABLES=1 keeps them in global memory (A/B measurements)
int csl_small_stage_bytes(const csl_program_t* p) {
  static const int off = csl_env_int("CSL_NO_STAGE_TABLES", 0);
  const int tb = csl_small_tables_bytes(p->ncode, p->nbox, p->ngauss);
  return (off || tb > CSL_SMALL_STAGE_MAX) ? 0 : tb;
}

// ---- tuning options (sweeps, A/B tests): process-wide integers, initialised from the environment variable
// CSL_<NAME IN CAPITALS> on first use; -1 = "let the launcher decide"
struct csl_option { const char* name; const char* env; int value; bool init; };
static csl_option g_options[] = {
    {"pipe", "CSL_PIPE", -1, false},                    // tile movement of the DMMA kernels: 0 cp.async, 1 bulk copies + mbarrier
    {"trigemm_rows", "CSL_TRIGEMM_ROWS", -1, false},    // 32 / 64: CTA shape of k_trigemm_chi2
    {"trigemm_group", "CSL_TRIGEMM_GROUP", -1, false},  // column tiles per super-group of k_trigemm_chi2
    {"trigemm_stages", "CSL_TRIGEMM_STAGES", -1, false},
    {"trsm_shape", "CSL_TRSM_SHAPE", -1, false},        // k_trsm_chi2 variant
    {"slab_warps", "CSL_SLAB_WARPS", -1, false},        // k_trsm_chi2_slab: all-wide CTAs of this many consumer warps
    {"bin_split", "CSL_BIN_SPLIT", -1, false},          // ell splits of k_bin_residual (0 = by launch size)
    {"bin_variant", "CSL_BIN_VARIANT", -1, false},      // k_bin_residual variant
    {"sab_threads", "CSL_SAB_THREADS", -1, false},      // CTA size of k_stretch_combine_accept
    {"seg_pair", "CSL_SEG_PAIR", -1, false},            // k_bin_segsum_pair: 0 = one launch per class
    {"seg_small", "CSL_SEG_SMALL", -1, false},          // k_bin_segsum: 1 walker per thread below this many walkers
};
static csl_option* find_option(const char* name) {
  for (csl_option& o : g_options)
    if (name && strcmp(o.name, name) == 0) {
      if (!o.init) { o.value = csl_env_int(o.env, -1); o.init = true; }
      return &o;
    }
  return nullptr;
}
int csl_option_value(const char* name) {
  csl_option* o = find_option(name);
  return o ? o->value : -1;
}
extern "C" int csl_set_option(const char* name, int value) {
  csl_option* o = find_option(name);
  if (!o) return csl_fail(CSL_E_BADARG, "csl_set_option: unknown option");
  o->value = value;
  return 0;
}
extern "C" int csl_get_option(const char* name) { return csl_option_value(name); }

// the four stages on W walkers whose workspace columns start at the given pointers and have stride ld
static int lnl_ld(const csl_program_t* p, int W, const double* x, double* lnl, double* ws_coef, double* ws_prior,
                  double* ws_R, double* ws_chi2, double* ws_part, int ld, void* stream) {
  const int Wp = round_up(W, CSL_BN);
  int rc = csl_eval_prepare_ld(p, W, p->ncoef > 0 ? Wp : W, x, ws_coef, ld, ws_prior, stream);
  if (rc) return rc;
  const double* chi2 = nullptr;
  if (p->resid_mode != 0) {
    CSL_REQUIRE(ws_R && ws_chi2 && ws_coef, "csl_lnl: workspace missing for the quadratic form");
    if (p->resid_mode == 1 || p->resid_mode == 3) rc = csl_bin_residual(p, Wp, ws_coef, ld, ws_R, ld, stream);
    else rc = csl_direct_residual_ld(p, W, Wp, ws_coef, ld, ws_R, ld, stream);
    if (rc) return rc;
    if (p->resid_mode == 3) rc = csl_cov_chi2(p, W, ws_R, ld, ws_chi2, stream);
    else if (p->chi2_mode == 1) {                     // the partial sums are added by the combine kernel
      rc = csl_chi2_parts(p, Wp, ws_R, ld, ws_part, stream);
      if (rc) return rc;
      return csl_combine_parts(p, W, ws_coef, ld, ws_prior, ws_part, ld, lnl, stream);
    } else rc = csl_chi2(p, Wp, ws_R, ld, ws_chi2, ws_part, stream);
    if (rc) return rc;
    chi2 = ws_chi2;
  }
  return csl_combine(p, W, ws_coef, ld, ws_prior, chi2, lnl, stream);
}

extern "C" int csl_lnl(const csl_program_t* p, int W, const double* x, double* lnl, double* ws_coef,
                       double* ws_prior, double* ws_R, double* ws_chi2, double* ws_part, void* stream) {
  CSL_REQUIRE(p && x && lnl && ws_prior && W > 0, "csl_lnl: null argument");
  return lnl_ld(p, W, x, lnl, ws_coef, ws_prior, ws_R, ws_chi2, ws_part, round_up(W, CSL_BN), stream);
}

// two library-owned side streams + events per device for the end-to-end call: the second half-batch's H2D copy
// and the first half-batch's D2H copy run under the other half's kernels
struct csl_side { cudaStream_t st[2]; cudaEvent_t fork, join[2]; bool ok; };
static csl_side* side_streams() {
  static csl_side sides[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  csl_side* sd = &sides[dev];
  if (!sd->ok) {
    for (int i = 0; i < 2; ++i) {
      if (cudaStreamCreateWithFlags(&sd->st[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&sd->join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    sd->ok = true;
  }
  return sd;
}

// The end-to-end call returns when its results are in host memory.  It waits by POLLING the stream: whether a blocking
// cudaStreamSynchronize spins or yields is the driver's choice, and a yielding wait wakes up 0.1-0.2 ms late -- a sixth
// of a 1.3 ms call.  Otherwise identical runs of the blocking version gave 49.9 M evaluations/s on most boxes and
// 42.6-43.6 M on some (profiles/r2_e2e_wait.json); where the blocking wait spins anyway the two are equal (49.9 both).
// CSL_E2E_BLOCKING=1 restores the blocking wait.
static cudaError_t wait_stream(cudaStream_t s) {
  static const int blocking = csl_env_int("CSL_E2E_BLOCKING", 0);
  if (blocking) return cudaStreamSynchronize(s);
  cudaError_t e;
  while ((e = cudaStreamQuery(s)) == cudaErrorNotReady) {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
  }
  return e;
}

extern "C" int csl_lnl_host(const csl_program_t* p, int W, const double* x_host, double* lnl_host, double* x_dev,
                            double* lnl_dev, double* ws_coef, double* ws_prior, double* ws_R, double* ws_chi2,
                            double* ws_part, void* stream) {
  CSL_REQUIRE(p && x_host && lnl_host && x_dev && lnl_dev && W > 0, "csl_lnl_host: null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const int ld = round_up(W, CSL_BN), nd = p->ndim;
  static const int single = csl_env_int("CSL_E2E_SINGLE", 0);
  csl_side* sd = (W >= 4 * CSL_BN && !single) ? side_streams() : nullptr;
  cudaError_t e;
#define CSL_CUDA(call) do { e = (call); if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e)); } while (0)
  if (!sd) {
    CSL_CUDA(cudaMemcpyAsync(x_dev, x_host, (size_t)W * nd * sizeof(double), cudaMemcpyHostToDevice, s));
    int rc = lnl_ld(p, W, x_dev, lnl_dev, ws_coef, ws_prior, ws_R, ws_chi2, ws_part, ld, stream);
    if (rc) return rc;
    CSL_CUDA(cudaMemcpyAsync(lnl_host, lnl_dev, (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, s));
    CSL_CUDA(wait_stream(s));
    return 0;
  }
  // halves are whole 128-walker tiles, so every walker's arithmetic is what the single batch does
  const int w_half = round_up((W + 1) / 2, CSL_BN);
  CSL_CUDA(cudaEventRecord(sd->fork, s));
  for (int c = 0; c < 2; ++c) {
    const int w0 = c * w_half, wc = c == 0 ? w_half : W - w_half;
    if (wc <= 0) continue;
    cudaStream_t sc = sd->st[c];
    CSL_CUDA(cudaStreamWaitEvent(sc, sd->fork, 0));
    CSL_CUDA(cudaMemcpyAsync(x_dev + (size_t)w0 * nd, x_host + (size_t)w0 * nd, (size_t)wc * nd * sizeof(double),
                             cudaMemcpyHostToDevice, sc));
    int rc = lnl_ld(p, wc, x_dev + (size_t)w0 * nd, lnl_dev + w0, ws_coef ? ws_coef + w0 : nullptr, ws_prior + w0,
                    ws_R ? ws_R + w0 : nullptr, ws_chi2 ? ws_chi2 + w0 : nullptr, ws_part ? ws_part + w0 : nullptr, ld, sc);
    if (rc) return rc;
    CSL_CUDA(cudaMemcpyAsync(lnl_host + w0, lnl_dev + w0, (size_t)wc * sizeof(double), cudaMemcpyDeviceToHost, sc));
    CSL_CUDA(cudaEventRecord(sd->join[c], sc));
    CSL_CUDA(cudaStreamWaitEvent(s, sd->join[c], 0));
  }
  CSL_CUDA(wait_stream(s));
#undef CSL_CUDA
  return 0;
}

// ---- host-released gate (bench.py: short timed regions start with their launches already queued) -------------
__global__ void k_host_gate(const volatile int32_t* flag, long long timeout_ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (*flag == 0) {
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if ((long long)(t - t0) > timeout_ns) break;
    __nanosleep(500);
  }
}
extern "C" int csl_host_gate(const int32_t* host_flag, double timeout_s, void* stream) {
  CSL_REQUIRE(host_flag && timeout_s > 0 && timeout_s <= 30.0, "csl_host_gate: null flag or timeout outside (0, 30] s");
  k_host_gate<<<1, 1, 0, (cudaStream_t)stream>>>(host_flag, (long long)(timeout_s * 1e9));
  return csl_check_launch("k_host_gate");
}

// ---- whole stretch-move steps enqueued from C -------------------------------------------------------------
extern "C" int csl_stretch_run(const csl_program_t* p, csl_ensemble_t* e, int nsteps, uint32_t step0, int64_t last_step,
                               void* stream) {
  CSL_REQUIRE(p && e && nsteps >= 0, "csl_stretch_run: null argument");
  CSL_REQUIRE(e->ndim == p->ndim && e->n0 >= 0 && e->n1 >= 0 && e->h0 > 0 && e->h1 > 0, "csl_stretch_run: bad ensemble");
  CSL_REQUIRE(e->pos && e->lnl && e->q && e->lnl_q && e->zz && e->weight && e->pos_old && e->rows && e->nrows,
              "csl_stretch_run: ensemble buffer missing");
  CSL_REQUIRE(e->replica == nullptr || (e->peer_ptrs && e->flag_ptrs && e->npeers > 0 && e->sync),
              "csl_stretch_run: replica given without peer / flag pointers / sync words");
  const int nd = e->ndim, W = e->n0 + e->n1;
  // binned bandpower likelihoods take the fused stages; everything else the staged ones
  static const bool no_fuse = csl_env_int("CSL_NO_FUSE", 0) != 0;
  const bool fused = p->resid_mode == 1 && e->ws_coef && e->ws_R && e->ws_chi2 && !no_fuse;
  // the row bookkeeping of a step (k_emcee_rows) rides in the accept kernel of each half: one launch fewer per step
  static const bool no_fold = csl_env_int("CSL_NO_FOLD_ROWS", 0) != 0;
  const bool fold_rows = fused && !no_fold;
  for (int i = 0; i < nsteps; ++i) {
    const uint32_t step = step0 + (uint32_t)i;
    for (int h = 0; h < 2; ++h) {
      const int ns = h == 0 ? e->n0 : e->n1, r0 = h == 0 ? 0 : e->n0;
      const int csize = h == 0 ? e->h1 : e->h0;
      const uint32_t g0 = h == 0 ? e->gid0 : e->gid1;
      // the complementary half: all walkers of the other colour (replica), or the local other slice
      const double* comp = e->replica ? e->replica + (long long)(h == 0 ? e->h0 : 0) * nd
                                      : e->pos + (long long)(h == 0 ? e->n0 : 0) * nd;
      int rc = 0;
      if (ns > 0) {
        double* s = e->pos + (long long)r0 * nd;
        double* q = e->q + (long long)r0 * nd;
        uint8_t* acc = e->accepted ? e->accepted + r0 : nullptr;
        if (fused) {
          // propose+prepare | residual | chi^2 (partials) | combine+accept(+publish+signal): 2 + nclass launches
          // per half.  Sharded: the proposal kernel WAITS for every peer's previous half-step (epoch), the accept
          // kernel's last CTA SIGNALS this one (epoch + 1): no separate barrier launch.
          const int Wp = round_up(ns, CSL_BN);
          csl_peer_args pa;
          pa.flag_ptrs = e->flag_ptrs; pa.sync = e->sync; pa.epoch = e->epoch; pa.rank = e->rank; pa.world = e->npeers;
          rc = csl_fused_propose_prepare(p, ns, Wp, csize, s, comp, e->a, e->seed, step, (uint32_t)h, g0, q,
                                         e->zz + r0, e->ws_coef, Wp, e->ws_prior, e->replica ? &pa : nullptr, stream);
          if (rc) return rc;
          rc = csl_bin_residual(p, Wp, e->ws_coef, Wp, e->ws_R, Wp, stream);
          if (rc) return rc;
          rc = csl_chi2_impl(p, Wp, e->ws_R, Wp, e->ws_chi2, e->ws_part, false, stream);
          if (rc) return rc;
          const bool parts = p->chi2_mode == 1;
          pa.epoch = e->epoch + 1;
          rc = csl_fused_combine_accept(p, ns, csize, nd, s, e->lnl + r0, q, e->lnl_q, e->zz + r0, e->seed, step,
                                        (uint32_t)h, g0, acc, e->naccepted, e->replica ? e->peer_ptrs : nullptr,
                                        e->replica ? e->npeers : 0, e->replica ? e->multicast_ptr : 0, (int64_t)g0,
                                        e->ws_coef, Wp, e->ws_prior, parts ? nullptr : e->ws_chi2,
                                        parts ? e->ws_part : nullptr, p->n_pad / 32, Wp, e->replica ? &pa : nullptr,
                                        fold_rows ? e->weight + r0 : nullptr,
                                        fold_rows ? e->rows + (long long)r0 * e->cap * (2 + nd) : nullptr,
                                        fold_rows ? e->nrows + r0 : nullptr, e->cap, (int64_t)step == last_step ? 1 : 0,
                                        stream);
          if (rc) return rc;
          if (e->replica) ++e->epoch;
          continue;
        }
        rc = csl_stretch_propose(ns, csize, nd, s, comp, e->a, nullptr, nullptr, e->seed, step, (uint32_t)h, g0, q,
                                 e->zz + r0, stream);
        if (rc) return rc;
        rc = csl_lnl(p, ns, q, e->lnl_q, e->ws_coef, e->ws_prior, e->ws_R, e->ws_chi2, e->ws_part, stream);
        if (rc) return rc;
        if (e->replica)
          rc = csl_stretch_accept_publish(ns, csize, nd, s, e->lnl + r0, q, e->lnl_q, e->zz + r0, nullptr, e->seed, step,
                                          (uint32_t)h, g0, acc, e->naccepted, e->peer_ptrs, e->npeers, e->multicast_ptr,
                                          (int64_t)g0, stream);
        else
          rc = csl_stretch_accept(ns, csize, nd, s, e->lnl + r0, q, e->lnl_q, e->zz + r0, nullptr, e->seed, step,
                                  (uint32_t)h, g0, acc, e->naccepted, stream);
        if (rc) return rc;
        if (e->replica) {
          rc = csl_peer_barrier(e->flag_ptrs, e->rank, e->npeers, ++e->epoch, e->sync, stream);
          if (rc) return rc;
        }
        continue;
      }
      if (e->replica) {       // no walker of this colour here: the peers still wait for this rank's signal
        if (fused) rc = csl_peer_signal(e->flag_ptrs, e->rank, e->npeers, ++e->epoch, stream);
        else rc = csl_peer_barrier(e->flag_ptrs, e->rank, e->npeers, ++e->epoch, e->sync, stream);
        if (rc) return rc;
      }
    }
    if (!fold_rows) {
      int rc = csl_emcee_rows(W, nd, e->pos_old, e->pos, e->lnl, e->weight, (int64_t)step == last_step ? 1 : 0, e->rows,
                              e->nrows, e->cap, stream);
      if (rc) return rc;
    }
  }
  // fused + sharded: the last half-step has been signalled but not waited for -- whoever reads the replica next
  // (a per-stage step, the host) must see every peer's stores
  if (e->replica && fused && nsteps > 0) {
    int rc = csl_peer_wait(e->flag_ptrs, e->rank, e->npeers, e->epoch, e->sync, stream);
    if (rc) return rc;
  }
  return 0;
}
