/*
 * cosmoslik_b200 -- C ABI of the B200 hot path (libcosmoslik_b200.so).
 *
 * The reference (marius311/cosmoslik) is pure Python and has no FFI; the
 * functions below are what a ctypes binding in the reference's plugins would
 * call for the path "sampler step + priors + foreground model + window binning
 * + Gaussian chi^2" (SURVEY.md section 8).  Each entry names the reference
 * lines it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - no function allocates, frees or synchronises (except *_host variants,
 *     which synchronise the given stream before returning);
 *   - `stream` is a cudaStream_t passed as void*; 0 is the legacy default stream;
 *   - return value 0 = success, otherwise a cudaError_t or a negative
 *     CSL_E_* code; csl_last_error() gives a message for the calling thread;
 *   - all floating point is IEEE binary64.
 *   - walker-major arrays are row-major [W, ndim]; "rows" matrices are
 *     [n_pad, ldr] with the WALKER index contiguous.
 */
#ifndef COSMOSLIK_B200_H
#define COSMOSLIK_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSL_ABI_VERSION 2

/* error codes (negative; positive values are cudaError_t) */
#define CSL_E_BADARG   (-1)
#define CSL_E_TOOLARGE (-2)
#define CSL_E_NODEVICE (-3)

/* geometry the host must honour when it pads its tables */
#define CSL_BM        64   /* rows per binning / TRSM row block            */
#define CSL_BN       128   /* walkers per column tile                      */
#define CSL_KC        16   /* ell (k) chunk                                */
#define CSL_MAX_DIM   64   /* max sampled parameters                       */
#define CSL_MAX_TERMS  8   /* template terms per spectrum                  */
#define CSL_MAX_PLAW   4   /* power-law terms per spectrum                 */
#define CSL_SPEC_STRIDE 64 /* int32 words per spectrum record              */
#define CSL_TILE_STRIDE 32 /* int32 words per tile record                  */
#define CSL_STACK     16   /* coefficient bytecode stack depth             */
#define CSL_MAX_CLASS  8   /* distinct (NT, NP) spectrum classes per program */
#define CSL_BIN_STRIDE 16  /* int32 words per bin record (segmented-sum binning) */
#define CSL_SEG_GROUP  8   /* bins per CTA row of the segmented-sum kernel    */
#define CSL_GRP_STRIDE 128 /* int32 words per group record (segmented-sum binning) */

/* coefficient bytecode (one flat program; STORE pops into coef row `arg`) */
enum {
  CSL_OP_PUSHP = 0, CSL_OP_PUSHC = 1, CSL_OP_ADD = 2, CSL_OP_SUB = 3, CSL_OP_MUL = 4,
  CSL_OP_DIV = 5, CSL_OP_NEG = 6, CSL_OP_EXP = 7, CSL_OP_LOG = 8, CSL_OP_POW = 9,
  CSL_OP_SQRT = 10, CSL_OP_STORE = 11, CSL_OP_SQR = 12
};

/* spectrum record, int32 words (table spec_tab[nspec][CSL_SPEC_STRIDE]) */
enum {
  CSL_S_NLPAD = 0,   /* padded number of ell columns (multiple of CSL_KC)       */
  CSL_S_NTERM = 1,   /* NT: template terms, padded to the class size {0,1,2,4,8}   */
  CSL_S_NPLAW = 2,   /* NP: power-law terms, padded to the class size {0,1,2,4}    */
  CSL_S_CAL   = 3,   /* coef row multiplying the data of this spectrum, -1 = 1.0 */
  CSL_S_CALMODE = 5, /* 0: r = cal*d - b (spt_lowl.py:132, mspec);  1: r = d - cal*b (SPTSZ_lowl.py:68) */
  CSL_S_ABERR = 6,   /* 1: aberration -- cl[l] *= 1 + G[l] (ln cl[l] - ln cl[l-1]) before binning (SPTSZ_lowl.py:78-79,
                        84-89), G[l] = 0.26 * 1.23e-3 / (ln ell_l - ln ell_{l-1}), G[0] = 0 (DMMA-binned spectra only) */
  CSL_S_ABTAB = 7,   /* offset (doubles) of G [nl_pad] in `templates`                                  */
  CSL_S_LTAB  = 4,   /* offset (doubles) of the ell-table in `templates`: nl_pad rows of
                        [T_0..T_{NT-1}, (logB_q, M_q, D_q, 0) q<NP], D_q[l] = logB_q[l+1] - logB_q[l];
                        cl[l] = sum_t coef[TERM_COEF[t]] T_t[l]
                              + sum_q coef[PLAW_AMP[q]] M_q[l] exp(coef[PLAW_IDX[q]] logB_q[l])
                        (padding terms point at an all-zero coefficient row / zero columns)  */
  CSL_S_TERM_COEF = 8, CSL_S_PLAW_AMP = 24, CSL_S_PLAW_IDX = 28,
  CSL_S_IDXBOUND = 32 /* [CSL_MAX_PLAW]: float32 bits, a bound on |coef[PLAW_IDX[q]]| the host expects of every walker
                         (the parameter's prior box, stored NEGATIVE = hard: outside it lnl is +inf anyway;
                         else +4 = soft: walkers outside it take exp per column): fixes the Taylor degree of
                         the power-law recurrence per group of bins */
};

/* tile record, int32 words (table tile_tab[ntile][CSL_TILE_STRIDE]) */
enum {
  CSL_T_ROW0 = 0,    /* first residual row of the tile                          */
  CSL_T_NROWS = 1,   /* valid rows, 1..CSL_BM                                   */
  CSL_T_SPEC = 2,
  CSL_T_LBEGIN = 3,  /* first ell column (multiple of CSL_KC)                   */
  CSL_T_LEND = 4,    /* one past the last column (multiple of CSL_KC)           */
  CSL_T_WINOFF_LO = 5, CSL_T_WINOFF_HI = 6, /* offset (doubles) of the tile's first window row   */
  CSL_T_WINLD = 7,   /* window row stride in doubles                            */
  CSL_T_GLO = 8,     /* [8]: per 8-row group, first non-zero column chunk start */
  CSL_T_GHI = 16,    /* [8]: per 8-row group, one past last non-zero column     */
  CSL_T_WINT_LO = 24, CSL_T_WINT_HI = 25  /* offset (doubles) of the tile's tile-major windows in winT (chunk 0) */
};

/* bin record, int32 words (table bin_tab[nsegbin][CSL_BIN_STRIDE]): one window row handled by the
 * segmented-sum kernel (block-sparse windows: rows whose supports barely overlap) */
enum {
  CSL_B_ROW = 0,     /* residual row                                            */
  CSL_B_SPEC = 1,
  CSL_B_LO = 2,      /* first non-zero column                                   */
  CSL_B_HI = 3,      /* one past the last non-zero column                       */
  CSL_B_WINOFF_LO = 4, CSL_B_WINOFF_HI = 5,  /* offset (doubles) of the window row */
  CSL_B_TABOFF_LO = 6, CSL_B_TABOFF_HI = 7,  /* offset (doubles, in `templates`) of the bin's weighted column rows
                                                [w T_0..w T_{NT-1}, w M_q.., D_q.., pad to an even count] for l in
                                                [lo, hi), NT = real template terms; the bins of a group are contiguous
                                                there, in record order */
  CSL_B_DMAX = 8     /* [CSL_MAX_PLAW]: float32 bits, max |D_q[l]| over [lo, hi) (see k_bin_segsum) */
};

/* group record, int32 words (table grp_tab[nsegbin / CSL_SEG_GROUP][CSL_GRP_STRIDE]): everything one CTA of
 * k_bin_segsum needs about its CSL_SEG_GROUP bins, precomputed by the host from the bin records so that the
 * kernel's prologue is one 512-byte load (record g describes bin records 8 g .. 8 g + 7) */
enum {
  CSL_G_PRE = 0,     /* [9]: columns of the group's bins laid back to back: bin b owns [PRE[b], PRE[b+1])   */
  CSL_G_LO = 9,      /* [8]: first ell column of each bin                                                   */
  CSL_G_ROW = 17,    /* [8]: residual row of each bin, -1 = padding                                         */
  CSL_G_CONT = 25,   /* [8]: 1 if the bin starts at the column where the previous bin of the group ended    */
  CSL_G_MODE = 33,   /* Taylor mode of the group: 0/1/2 = degree 5/6/10, 3 = exp at every column            */
  CSL_G_BASE_LO = 34, CSL_G_BASE_HI = 35,   /* offset (doubles, in `templates`) of the group's weighted columns */
  CSL_G_SPEC = 36,
  CSL_G_DATA = 40,   /* [8] doubles: data value of each bin's residual row                                  */
  CSL_G_LOGB = 56    /* [8][CSL_MAX_PLAW] doubles: logB_q at each bin's first column                        */
};

/*
 * A compiled posterior ("device program").  Built on the host from the plugin
 * tree; the struct itself lives in host memory, its pointers are device
 * pointers.  It states, for W walkers at once, what Slik.evaluate computes
 * for one (cosmoslik/cosmoslik.py:109-141):
 *
 *   prior  = likelihoods/priors.py:24-30   (box -> +inf, Gaussian penalties); box_idx / gauss_idx >= 0 name a
 *            sampled parameter, < 0 a DERIVED quantity: coefficient row -idx - 1 (cosmoslik.py:139-141: priors
 *            are evaluated again after the script's __call__ so that they can depend on what it sets)
 *   coef   = scalar functions of the parameters (bytecode)
 *   lnl    = prior + sum_k coef[scalar_coef[k]] + 1/2 || L^-1 r ||^2
 *   r      = cal * d - W . cl        (resid_mode 1; spt_lowl.py:132,171-182;
 *                                     mspec_lnl.py:63-73)
 *          = coef[direct_coef[i]] - d[i]   (resid_mode 2; dense Gaussian)
 *   resid_mode 3: r = d - cal * W.cl and V_k = W_k.cl, lnl = 1/2 r^T (Sigma + V V^T)^-1 r
 *   cl     = sum_t coef * T_t[l] + sum_p coef * M_p[l] * exp(coef * logB_p[l])
 *            (spt_lowl.py:206-207; models/clust_poisson_egfs.py:14-17)
 */
typedef struct csl_program {
  int32_t ndim, ncoef, ncode, nbox, ngauss, nscalar, nspec, ntile;
  int32_t n, n_pad, resid_mode, nclass;
  int32_t cls[CSL_MAX_CLASS][4];   /* per class: NT, NP, first tile, number of tiles (tiles sorted by class) */
  const int32_t* code_op;   const int32_t* code_arg;  const double* code_val;
  const int32_t* box_idx;   const double* box_lo;     const double* box_hi;
  const int32_t* gauss_idx; const double* gauss_c;    const double* gauss_w;
  const int32_t* scalar_coef;
  const int32_t* spec_tab;  const int32_t* tile_tab;
  const double* templates;  const double* windows;    const double* data;
  const int32_t* direct_coef;
  const double* Lmat;       /* [n_pad, n_pad] row-major lower factor, identity padded */
  const double* Linv_diag;  /* [n_pad/CSL_BM][CSL_BM*CSL_BM] inverses of the diagonal blocks */
  const double* Minv;       /* [n_pad, n_pad] row-major L^-1 (lower), or NULL                     */
  int32_t chi2_mode;        /* 0: blocked forward substitution; 1: explicit-inverse tile product  */
  int32_t nsegclass;        /* segmented-sum classes (spectra with block-sparse windows)          */
  int32_t segcls[CSL_MAX_CLASS][4];  /* per class: NT (the spectra's REAL number of template terms, 0..8), NP (padded:
                                        0/1/2/4), first bin record, number of bin records (a multiple of CSL_SEG_GROUP) */
  const int32_t* bin_tab;
  /* resid_mode 3: model-dependent covariance C = Sigma + V V^T (SPTSZ_lowl.py:62-81).  The residual
   * vector stacks r (n_base rows) and the nmode columns of V (n_base rows each); chi^2 = r^T C^-1 r by a
   * per-walker Cholesky factorisation (csl_cov_chi2). */
  const double* Sigma;      /* [n_base, n_base] row-major */
  int32_t n_base, nmode;
  const int32_t* grp_tab;   /* group records of the segmented-sum bins (CSL_G_*), one per CSL_SEG_GROUP bin records */
  /* TILE-MAJOR copies of the A operands for the bulk-copy (TMA) pipelines: tile (rb, c) = rows 64 rb .. 64 rb + 63,
   * columns 16 c .. 16 c + 15 of the matrix, stored as 64 rows of 16 + 4 (padding) doubles = 1280 contiguous
   * doubles at offset (rb * (ncols / 16) + c) * 1280, so that ONE cp.async.bulk moves a tile into its
   * conflict-free shared-memory layout.  NULL: the kernels fall back to the row-major tables and cp.async. */
  const double* MinvT;      /* of Minv      [n_pad, n_pad]                                            */
  const double* LmatT;      /* of Lmat      [n_pad, n_pad]                                            */
  const double* LinvT;      /* of Linv_diag [n_pad/64][64, 64]: tile (b, c), c < 4                    */
  const double* winT;       /* of windows: per tile record, CSL_T_WINT_LO/HI = offset (doubles) of the tile's
                               chunk 0; chunk c of the tile follows at + c * 1280                      */
  int32_t bin_split;        /* 1, 2 or 4: ell split of the DMMA binning tiles (thread-block cluster size; the partial
                               sums are added in rank order).  Fixed per PROGRAM by the host from the chunks per
                               tile -- never from the launch size -- so that a walker's result does not depend on
                               the batch it is evaluated in.                                            */
  int32_t has_aberr;        /* some spectrum carries CSL_S_ABERR                                          */
} csl_program_t;

int csl_abi_version(void);
const char* csl_last_error(void);
int csl_sizeof_program(void);
int csl_sizeof_ensemble(void);
/* device properties: out[0]=SM count, out[1]=cc major, out[2]=cc minor, out[3]=max smem/block optin */
int csl_device_info(int device, int32_t* out4_host);

/* ---- posterior evaluation (replaces Slik.evaluate, cosmoslik.py:109-141) ---------------- */

/* priors (likelihoods/priors.py:24-30) + coefficient bytecode.
 * x [W, ndim] -> coef [ncoef, ldc] (ldc >= W), prior [W] (+inf outside a box).
 * Coefficient columns W..ldc-1 are filled with walker W-1's values (tile padding). */
int csl_eval_prepare(const csl_program_t* p, int W, const double* x, double* coef, int ldc,
                     double* prior, void* stream);

/* foreground build in shared memory (spt_lowl.py:206-207, clust_poisson_egfs.py:14-17),
 * window binning (spt_lowl.py:182) and the residual r = cal*d - b (spt_lowl.py:132) -> R [n_pad, ldr].
 * Spectra with dense windows go through the FP64 DMMA tile product (tile_tab / cls); spectra whose
 * window rows barely overlap (top-hat bins) through a per-bin segmented sum (bin_tab / segcls) that
 * builds every cl[l] exactly once.  W must be a multiple of CSL_BN (pad with copies of a valid walker). */
int csl_bin_residual(const csl_program_t* p, int W, const double* coef, int ldc, double* R, int ldr,
                     void* stream);
/* number of kernel launches one csl_bin_residual call makes for this program (one per spectrum class; two
 * segmented-sum classes whose pair has a fused kernel go out as one launch).  No reference counterpart. */
int csl_bin_residual_launches(const csl_program_t* p);

/* resid_mode 2: R[i, w] = coef[direct_coef[i], w] - data[i] (zero rows up to n_pad). */
int csl_direct_residual(const csl_program_t* p, int W, const double* coef, int ldc, double* R, int ldr,
                        void* stream);

/* chi2[w] = || L^-1 R[:, w] ||^2 on DMMA (replaces cho_solve + dot, spt_lowl.py:133 / mspec_lnl.py:73).
 *   chi2_mode 0: blocked forward substitution, 64-row blocks, R is overwritten by L^-1 R;
 *   chi2_mode 1: tile products with the host-inverted factor Minv (chosen by the host only when the
 *                inverse reproduces the substitution result to 1e-13); needs ws_part [n_pad/32, ldr] (one partial sum per 32 rows). */
int csl_chi2(const csl_program_t* p, int W, double* R, int ldr, double* chi2, double* ws_part, void* stream);
/* Explicit-inverse programs only: the tile product alone, the column sums of squares of every 32 rows of Y = L^-1 R
 * left in ws_part[n_pad/32, ldr] for csl_combine_parts (one reduction launch fewer than csl_chi2 + csl_combine). */
int csl_chi2_parts(const csl_program_t* p, int W, double* R, int ldr, double* ws_part, void* stream);

/* resid_mode 3 (replaces cho_factor(beam_cov + sigma) + cho_solve per call, SPTSZ_lowl.py:66-69):
 * R [(1 + nmode) * n_base (padded), ldr] holds r and the columns of V for every walker; one warp per
 * walker builds C = Sigma + V V^T in shared memory, factors it and solves.  n_base <= 64, nmode <= 32. */
int csl_cov_chi2(const csl_program_t* p, int W, const double* R, int ldr, double* chi2, void* stream);

/* lnl[w] = prior[w] + sum scalar coefs + chi2[w]/2 ; +inf where prior is +inf (cosmoslik.py:135-141).
 * chi2 may be NULL when the program has no quadratic form. */
int csl_combine(const csl_program_t* p, int W, const double* coef, int ldc, const double* prior,
                const double* chi2, double* lnl, void* stream);
/* The same for a program in explicit-inverse mode whose chi^2 is still in the per-32-row partial sums csl_chi2_parts
 * left in part[n_pad/32, ldp]: they are added here in row order (the order csl_chi2 adds them), then combined. */
int csl_combine_parts(const csl_program_t* p, int W, const double* coef, int ldc, const double* prior,
                      const double* part, int ldp, double* lnl, void* stream);

/* The four calls above in sequence.  Workspace (device, caller owned):
 *   coef [ncoef, Wp], prior [Wp], R [n_pad, Wp], chi2 [Wp], part [n_pad/32, Wp]
 *   with Wp = W rounded up to CSL_BN. */
int csl_lnl(const csl_program_t* p, int W, const double* x, double* lnl,
            double* ws_coef, double* ws_prior, double* ws_R, double* ws_chi2, double* ws_part, void* stream);

/* End-to-end variant with HOST buffers (pinned or pageable): copies x_host -> x_dev, evaluates,
 * copies lnl back and synchronises `stream`.  x_dev [Wp, ndim] and lnl_dev [Wp] are device scratch. */
int csl_lnl_host(const csl_program_t* p, int W, const double* x_host, double* lnl_host,
                 double* x_dev, double* lnl_dev,
                 double* ws_coef, double* ws_prior, double* ws_R, double* ws_chi2, double* ws_part,
                 void* stream);

/* ---- Metropolis-Hastings (samplers/metropolis_hastings.py) -------------------------------- */

/* draw_x (:141-142) for all chains: test_x = cur_x + delta.
 *   mode 0: delta = inj[w, :]                     (injected offsets; exact add)
 *   mode 1: delta = inj[w, :] @ F                 (injected standard normals)
 *   mode 2: delta = philox_normals(seed, step, chain0 + w) @ F
 * F [ndim, ndim] row-major is the matrix numpy's multivariate_normal uses (host SVD). */
int csl_mh_propose(int W, int ndim, const double* cur_x, double* test_x, int mode, const double* inj,
                   const double* F, uint64_t seed, uint32_t step, uint32_t chain0, void* stream);

/* accept test + weight bookkeeping + row emission (:154-164) for all chains.
 *   u: injected uniforms [W] or NULL -> Philox(seed, step, chain0 + w).
 *   rows [W, cap, 2+ndim] receives (lnl, weight, x...) at slot nrows[w]++ when the reference
 *   would yield; accepted [W] (uint8) records the decision, margin [W] = log(u) - temp*(cur-test)
 *   (for near-tie diagnostics; may be NULL). */
int csl_mh_accept(int W, int ndim, double* cur_x, double* cur_lnl, double* cur_weight,
                  const double* test_x, const double* test_lnl, const double* u,
                  uint64_t seed, uint32_t step, uint32_t chain0, double temp, double max_weight,
                  double* rows, int32_t* nrows, int cap, uint8_t* accepted, double* margin, void* stream);

/* Whole MH run for a dense Gaussian posterior, one warp per chain, `nsteps` steps in ONE launch
 * (draw_x + priors + x^T C^-1 x/2 by forward substitution + accept + rows); ndim <= 32.
 *   Lfac: LOWER Cholesky factor [n, n] of the target covariance, mean [n]; residual row i is
 *   x[perm[i]] - mean[i].  Priors are taken from p.
 *   z/u: injected streams [nsteps, W, ndim] / [nsteps, W], or NULL for Philox.
 *   z_is_delta: 1 -> z holds offsets. */
int csl_mh_gauss_run(const csl_program_t* p, int W, int nsteps, const double* Lfac, const double* mean,
                     const int32_t* perm, double* cur_x, double* cur_lnl, double* cur_weight,
                     const double* z, int z_is_delta, const double* u, const double* F,
                     uint64_t seed, uint32_t step0, uint32_t chain0, double temp, double max_weight,
                     double* rows, int32_t* nrows, int cap, void* stream);

/* State of one rank's share of a batch of MH chains, for csl_mh_run (all pointers device pointers, caller owned). */
typedef struct csl_mh {
  int32_t ndim, W, cap;          /* parameters, chains on this GPU, row-buffer capacity per chain              */
  uint32_t chain0;               /* global id of local chain 0 (Philox key)                                   */
  uint64_t seed;
  double temp, max_weight;       /* metropolis_hastings.py:43, :41                                            */
  int32_t yield_rejected;        /* :160 -- a rejected proposal is emitted as a weight-0 row                   */
  int32_t inj_is_delta;          /* inj_z holds offsets (exact add) instead of standard normals               */
  double *cur_x, *cur_lnl, *cur_weight, *test_x, *test_lnl;
  double* rows;                  /* [W, cap, 2 + ndim]: (lnl, weight, x...) at slot nrows[w]++                */
  int32_t* nrows;
  uint8_t* accepted;             /* [W] last decision, or NULL                                                */
  double* margin;                /* [W] log(u) - temp (cur - test) of the last step, or NULL                  */
  uint64_t* naccepted;           /* device counter, or NULL                                                   */
  const double* F;               /* [ndim, ndim] row-major: test_x = cur_x + z F (rewritten in place by
                                    csl_mh_adapt_factor between blocks)                                       */
  uint32_t* step;                /* DEVICE step counter: Philox step number of the next MH step; csl_mh_run
                                    advances it by nsteps (a block's kernel arguments never change, so one CUDA
                                    graph serves every block)                                                 */
  const double* inj_z;           /* injected streams for the steps of ONE call: [nsteps, W, ndim], or NULL     */
  const double* inj_u;           /* [nsteps, W], or NULL -> Philox                                            */
  double *ws_coef, *ws_prior, *ws_R, *ws_chi2, *ws_part;   /* csl_lnl workspace for W chains                  */
} csl_mh_t;
int csl_sizeof_mh(void);

/* `nsteps` MH steps of every chain (replaces the loop of _mcmc, :150-164, for any compiled posterior): per step
 * the fused proposal + priors + bytecode kernel, the residual and chi^2 kernels of the program, and the fused
 * lnl-combine + accept + weight + row-emission kernel, enqueued on `stream` from C.  Chain contents are bit for
 * bit those of csl_mh_propose / csl_lnl / csl_mh_accept. */
int csl_mh_run(const csl_program_t* p, const csl_mh_t* m, int nsteps, void* stream);

/* The same block of `nsteps` steps captured ONCE into a CUDA graph (an internal capture stream is used: the
 * caller's stream may be the legacy default stream, which cannot capture); csl_graph_launch replays it on
 * `stream` -- one host call per block.  The state's pointers and nsteps are frozen into the graph; the step
 * counter, the proposal factor and the injected-stream buffers are read through their pointers at run time.
 * The graph handle is host memory owned by the library until csl_graph_destroy. */
int csl_mh_graph_create(const csl_program_t* p, const csl_mh_t* m, int nsteps, void** graph_out);
int csl_graph_launch(void* graph, void* stream);
int csl_graph_nodes(void* graph);
int csl_graph_destroy(void* graph);

/* Proposal adaptation without leaving the device (replaces the master's get_new_cov + the SVD inside draw_x).
 * csl_mh_moments_fused: ONE pass over the last half of every chain's rows,
 *   out = [sum w, sum w (x - ref), sum w (x - ref)(x - ref)^T]   (1 + ndim + ndim^2 doubles; partial [592, that]);
 *   `ref` [ndim] is a fixed point (the start point), identical on every rank, so ONE all-reduce of `out` pools
 *   the ranks.
 * csl_mh_adapt_factor: from the (all-reduced) sums, cov_out = pooled covariance (denominator sum w - 1,
 *   chains.py:507-512), F_out = factor with F^T F = cov / ndim * proposal_scale^2 (Cholesky; a direction nobody
 *   moved in gives a zero column), mean_out (may be NULL) = pooled mean. */
int csl_mh_moments_fused(int W, int ndim, const double* rows, const int32_t* nrows, int cap, const double* ref,
                         double* partial, double* out, void* stream);
int csl_mh_adapt_factor(int ndim, const double* mom, double proposal_scale, const double* ref, double* cov_out,
                        double* F_out, double* mean_out, void* stream);

/* get_new_cov (:249-253) sufficient statistics over the last half of every chain's rows:
 * pass 1 (mean == NULL): out[0] = sum w, out[1..ndim] = sum w x.
 * pass 2 (mean given):   out[1+ndim ..] = sum w (x-mean)(x-mean)^T  [ndim*ndim].
 * partial [592, 1+ndim+ndim*ndim] is scratch; the reduction order is fixed (bitwise reproducible).
 * Sums over ranks are one all-reduce of out. */
int csl_mh_moments(int W, int ndim, const double* rows, const int32_t* nrows, int cap,
                   const double* mean, double* partial, double* out, void* stream);

/* The same two-pass weighted moments over ALL chains' rows from row `burn_rows` on: the device side of
 * Chain.mean / Chain.cov on the joined chains (chains.py:78-96, 345-355) without a host round trip. */
int csl_rows_moments(int W, int ndim, const double* rows, const int32_t* nrows, int cap, int burn_rows,
                     const double* mean, double* partial, double* out, void* stream);

/* Chain.add_gauss_prior on the device rows (chains.py:206-233): for the np parameters pidx [np] (device int32),
 * mean [np], Cinv [np, np] (device; one parameter: 1 / std^2), every stored row gets
 *   dlnl = (x_P - mean)^T Cinv (x_P - mean) / 2,  weight *= exp(-dlnl),  lnl += dlnl.   In place. */
int csl_rows_gauss_prior(int W, int ndim, double* rows, const int32_t* nrows, int cap, int np,
                         const int32_t* pidx, const double* mean, const double* cinv, void* stream);

/* Chain.thin(delta) on the device rows (chains.py:235-243): per chain, with S_k = ceil(cumulative weight through
 * row k / delta), row k is kept iff S_k > S_{k-1} and gets weight S_k - S_{k-1}; kept rows are compacted in place
 * and nrows updated. */
int csl_rows_thin(int W, int ndim, double* rows, int32_t* nrows, int cap, double delta, void* stream);

/* ---- affine-invariant ensemble sampler (samplers/emcee.py + emcee 2.x stretch move) ------- */

/* proposals for the Ns active walkers s [Ns, ndim] against the complementary half comp [Nc, ndim]:
 *   z = ((a-1) u_z + 1)^2 / a ;  q = c_j - z (c_j - s)
 * u_z [Ns], j [Ns] injected, or NULL -> Philox(seed, step, half, walker0 + k) with j over Nc.
 * Writes q [Ns, ndim] and zz [Ns]. */
int csl_stretch_propose(int Ns, int Nc, int ndim, const double* s, const double* comp, double a,
                        const double* u_z, const int64_t* j, uint64_t seed, uint32_t step, uint32_t half,
                        uint32_t walker0, double* q, double* zz, void* stream);

/* accept iff (ndim-1) ln z + lnp(q) - lnp(s) > ln u_acc, lnp = -lnl; updates s, lnl_s in place.
 * accepted [Ns] (uint8) and *naccepted (device counter, incremented by the number of accepted moves)
 * may be NULL. */
int csl_stretch_accept(int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q, const double* lnl_q,
                       const double* zz, const double* u_acc, uint64_t seed, uint32_t step, uint32_t half,
                       uint32_t walker0, uint8_t* accepted, uint64_t* naccepted, void* stream);

/* Accept fused with the exchange step of the sharded ensemble (replaces the mpi_map pool round trip,
 * mpi/mpi.py:40-64): besides updating s / lnl_s, every walker's final position is stored into row
 * dst_row0 + k of EVERY rank's replica of the ensemble -- peer_ptrs[npeers] are the device addresses of
 * the replicas (peer-mapped symmetric memory, this rank included) or, if multicast_ptr != 0, one
 * multimem.st through the NVSwitch multicast address.  The caller issues a device-side barrier on the
 * stream afterwards. */
int csl_stretch_accept_publish(int Ns, int Nc, int ndim, double* s, double* lnl_s, const double* q,
                               const double* lnl_q, const double* zz, const double* u_acc, uint64_t seed,
                               uint32_t step, uint32_t half, uint32_t walker0, uint8_t* accepted,
                               uint64_t* naccepted, const uint64_t* peer_ptrs, int npeers, uint64_t multicast_ptr,
                               int64_t dst_row0, void* stream);

/* Device-side barrier over the ranks of one NVLink domain, on `stream` (the host does not block).
 * flag_ptrs[world]: device addresses of every rank's uint64[world] flag array (symmetric memory, zeroed
 * once); epoch must increase by one per call.  Orders the peer stores of the preceding kernel.
 * sync: device uint32[2] owned by the caller, zeroed once -- sync[0] is a STICKY status word: a wait that
 * exceeds CSL_BARRIER_TIMEOUT_S (environment, default 120 s) sets bit 0 and every later wait returns at once, so
 * the host must read it before trusting results (samplers.emcee does, per output block); sync == NULL makes a
 * timed-out wait trap instead.  A timed-out barrier never passes silently.
 * Used by the per-stage path; csl_stretch_run folds the same signal / wait into its accept / proposal kernels. */
int csl_peer_barrier(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, uint32_t* sync, void* stream);
/* wait only: returns (stream ordered) once every peer has announced `epoch` */
int csl_peer_wait(const uint64_t* flag_ptrs, int rank, int world, uint64_t epoch, uint32_t* sync, void* stream);

/* Host-released gate on `stream`: a one-thread kernel that spins until *host_flag (an int32 in PINNED host memory,
 * read through its unified address) becomes non-zero, or timeout_s seconds have passed.  Work enqueued behind it
 * starts the moment the host writes the flag, with all of its launches already queued -- what a sampler run of
 * thousands of steps looks like to the device anyway (the host enqueues a step several times faster than the
 * device runs it).  bench.py opens short timed regions this way so that a host hiccup at the start of a 20-step
 * region is not booked as step time; nothing in the samplers uses it.  (No reference counterpart.) */
int csl_host_gate(const int32_t* host_flag, double timeout_s, void* stream);

/* wrapper bookkeeping of samplers/emcee.py:70-88 after a full step: a walker that did not move
 * gains weight; one that moved (or any walker on the last step) emits (lnl_next, weight, x_old) and its
 * pos_old row is refreshed to pos_new (ready for the next step). */
int csl_emcee_rows(int W, int ndim, double* pos_old, const double* pos_new, const double* lnl_new,
                   double* weight, int last_step, double* rows, int32_t* nrows, int cap, void* stream);

/* State of one rank's share of a stretch-move ensemble, for csl_stretch_run.  Local walkers are stored
 * first-half slice then second-half slice: rows [0, n0) and [n0, n0+n1) of pos / lnl / q / zz / weight. */
typedef struct csl_ensemble {
  int32_t ndim, n0, n1;          /* local walkers in the first / second half                          */
  int32_t h0, h1;                /* global sizes of the two halves                                      */
  uint32_t gid0, gid1;           /* global walker id of local row 0 and of local row n0                 */
  int32_t cap;                   /* row-buffer capacity per walker                                      */
  double a;                      /* stretch scale (emcee default 2)                                     */
  uint64_t seed;                 /* Philox key                                                          */
  double *pos, *lnl, *q, *lnl_q, *zz, *weight, *pos_old, *rows;
  int32_t* nrows;
  uint8_t* accepted;
  uint64_t* naccepted;
  double *ws_coef, *ws_prior, *ws_R, *ws_chi2, *ws_part;     /* csl_lnl workspace for max(n0, n1) walkers */
  /* sharded ensembles: replica of ALL positions in symmetric memory (NULL on a single GPU) */
  double* replica;               /* this rank's replica [h0 + h1, ndim]                                 */
  const uint64_t* peer_ptrs;     /* device array [npeers] of every rank's replica address               */
  const uint64_t* flag_ptrs;     /* device array [npeers] of every rank's barrier flags (uint64[npeers]) */
  uint64_t multicast_ptr;        /* NVSwitch multicast address of the replica, or 0                     */
  uint64_t epoch;                /* barrier epoch, advanced by csl_stretch_run                          */
  int32_t npeers, rank;
  uint32_t* sync;                /* device uint32[2], zeroed once: [0] sticky status (bit 0 = a peer wait timed out:
                                    the results of the run are void), [1] ticket counter of the publishing kernel;
                                    required when replica != NULL                                         */
} csl_ensemble_t;

/* `nsteps` full stretch-move steps (both half-updates: propose, priors, foregrounds, binning, chi^2,
 * accept [+ publish + peer barrier], then the wrapper's row bookkeeping) enqueued on `stream` from C --
 * one call replaces ~20 Python-level launches per step (the loop of samplers/emcee.py:66-88 around
 * emcee's EnsembleSampler.sample).  Steps are numbered step0, step0+1, ...; the step whose number equals
 * last_step emits every walker's row (emcee.py:70).  Device Philox streams only. */
int csl_stretch_run(const csl_program_t* p, csl_ensemble_t* e, int nsteps, uint32_t step0, int64_t last_step,
                    void* stream);

/* ---- convergence (new; nothing in the reference) ----------------------------------------- */

/* per-chain weighted count / mean / variance over rows -> stats [W, 1+2*ndim] */
int csl_chain_moments(int W, int ndim, const double* rows, const int32_t* nrows, int cap, int burn_rows,
                      double* stats, void* stream);

/* ---- tuning / test hooks ------------------------------------------------------------------ */
/* Process-wide integer options for sweeps and A/B tests (-1 = let the launcher decide; each is initialised from
 * the environment variable CSL_<NAME>): "pipe" (0 = cp.async pipelines, 1 = bulk copies + mbarrier),
 * "trigemm_rows", "trigemm_group", "trigemm_stages", "trsm_shape", "bin_split", "bin_variant", "sab_threads",
 * "seg_small".  Results never depend on them bit-wise unless stated in DESIGN.md. */
int csl_set_option(const char* name, int value);
int csl_get_option(const char* name);

int csl_philox_mh(int W, int ndim, uint64_t seed, uint32_t step, uint32_t chain0, double* z, double* u,
                  void* stream);
int csl_philox_stretch(int Ns, int Nc, uint64_t seed, uint32_t step, uint32_t half, uint32_t walker0,
                       double* u_z, int64_t* j, double* u_acc, void* stream);
/* plain FP64 DMMA GEMM C[M,N] = A[M,K] B[K,N] (row-major, M%64==0, N%128==0, K%16==0) for peak tests */
int csl_dgemm_dmma(int M, int N, int K, const double* A, const double* B, double* C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COSMOSLIK_B200_H */
