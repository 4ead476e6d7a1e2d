// chi^2 against a MODEL-DEPENDENT covariance  C_w = Sigma + V_w V_w^T  (one per walker).
//   reference: likelihoods/SPTSZ_lowl_2017/SPTSZ_lowl.py:62-81 -- beam_cov = W (B o dl dl^T) W^T with
//   B = sum_m b_m b_m^T, i.e. beam_cov = sum_m (W (b_m o dl)) (W (b_m o dl))^T = V V^T; the reference
//   calls cho_factor(beam_cov + sigma) and cho_solve on every evaluation (:66-69), no log-determinant.
// Here the columns V_k = W_k . cl come out of the binning kernel as extra residual rows (windows
// W_k = W diag(b_k) are prepared by the host); one warp per walker then builds C in shared memory
// (packed lower triangle), factors it in place (right-looking Cholesky, lanes over the trailing update)
// and forward-substitutes r.  A CTA of 4 warps takes 4 ADJACENT walkers so the loads of their residual
// rows are full 32-byte sectors.
#include "common.cuh"

#define COV_WARPS 4

__global__ void __launch_bounds__(COV_WARPS * 32)
k_cov_chol_chi2(csl_program_t p, int W, const double* __restrict__ R, int ldr, double* __restrict__ chi2) {
  extern __shared__ double sm[];
  const int n = p.n_base, K = p.nmode;
  const int ntri = n * (n + 1) / 2;
  const int per_warp = ntri + n * K + n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int w0 = blockIdx.x * COV_WARPS;
  // ---- cooperative load of r and V for the CTA's walkers
  for (int idx = tid; idx < (K + 1) * n * COV_WARPS; idx += COV_WARPS * 32) {
    const int row = idx / COV_WARPS, j = idx - row * COV_WARPS;
    const int w = min(w0 + j, W - 1);
    const double v = R[(long long)row * ldr + w];
    double* base = sm + j * per_warp;
    if (row < n) base[ntri + n * K + row] = v;                       // r
    else { const int k = row / n - 1, i = row - (k + 1) * n; base[ntri + i * K + k] = v; }   // V[i][k]
  }
  __syncthreads();
  double* C = sm + warp * per_warp;       // packed lower triangle, row-major: C[i(i+1)/2 + j], j <= i
  double* V = C + ntri;
  double* r = V + n * K;
  // ---- C = Sigma + V V^T
  for (int i = 0; i < n; ++i) {
    const double* vi = V + i * K;
    for (int j = lane; j <= i; j += 32) {
      const double* vj = V + j * K;
      double c = __ldg(p.Sigma + i * n + j);
      for (int k = 0; k < K; ++k) c = fma(vi[k], vj[k], c);
      C[i * (i + 1) / 2 + j] = c;
    }
  }
  __syncwarp();
  // ---- in-place Cholesky (lower), right-looking
  for (int j = 0; j < n; ++j) {
    const double d = sqrt(C[j * (j + 1) / 2 + j]);
    __syncwarp();
    if (lane == 0) C[j * (j + 1) / 2 + j] = d;
    const double inv = 1.0 / d;
    for (int i = j + 1 + lane; i < n; i += 32) C[i * (i + 1) / 2 + j] *= inv;
    __syncwarp();
    for (int i = j + 1; i < n; ++i) {
      const double lij = C[i * (i + 1) / 2 + j];
      for (int k = j + 1 + lane; k <= i; k += 32) C[i * (i + 1) / 2 + k] = fma(-lij, C[k * (k + 1) / 2 + j], C[i * (i + 1) / 2 + k]);
    }
    __syncwarp();
  }
  // ---- forward substitution L y = r, chi2 = |y|^2
  double ss = 0.0;
  for (int j = 0; j < n; ++j) {
    const double y = r[j] / C[j * (j + 1) / 2 + j];
    __syncwarp();
    for (int i = j + 1 + lane; i < n; i += 32) r[i] = fma(-C[i * (i + 1) / 2 + j], y, r[i]);
    __syncwarp();
    ss = fma(y, y, ss);
  }
  if (lane == 0 && w0 + warp < W) chi2[w0 + warp] = ss;
}

extern "C" int csl_cov_chi2(const csl_program_t* p, int W, const double* R, int ldr, double* chi2, void* stream) {
  CSL_REQUIRE(p && R && chi2 && W > 0 && ldr >= W, "csl_cov_chi2: bad argument");
  CSL_REQUIRE(p->resid_mode == 3 && p->Sigma && p->n_base >= 1 && p->n_base <= 64 && p->nmode >= 0 && p->nmode <= 32,
              "csl_cov_chi2: needs resid_mode 3, n_base <= 64, nmode <= 32");
  const int n = p->n_base, K = p->nmode;
  const size_t smem = (size_t)COV_WARPS * (n * (n + 1) / 2 + n * K + n) * sizeof(double);
  if (smem > 48 * 1024) {        // per device and cheap: set whenever the launch needs it
    cudaError_t e = cudaFuncSetAttribute(k_cov_chol_chi2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return csl_fail((int)e, cudaGetErrorString(e));
  }
  k_cov_chol_chi2<<<(W + COV_WARPS - 1) / COV_WARPS, COV_WARPS * 32, smem, (cudaStream_t)stream>>>(*p, W, R, ldr, chi2);
  return csl_check_launch("k_cov_chol_chi2");
}
