// Microbenchmark: do DMMA.8x8x4 and DFMA share an execution pipe on sm_100a?
// mode 0: all warps DMMA; mode 1: all warps DFMA; mode 2: even warps DMMA, odd warps DFMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void k(int mode, int iters, double* out) {
  int warp = threadIdx.x >> 5;
  bool do_mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
  double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
  double c[16];
  for (int i = 0; i < 16; ++i) c[i] = i;
  if (do_mma) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma(c[2 * i], c[2 * i + 1], a, b);
  } else {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0;
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int mode = 0; mode < 3; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k<<<148 * 4, 256>>>(mode, iters, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      // per warp: mode0: iters*8 DMMA (512 flop each); mode1: iters*128 DFMA warp-instr (64 flop each)
      double nw = 148.0 * 4 * 8;
      double fl_mma = (mode == 0 ? nw : mode == 2 ? nw / 2 : 0) * iters * 8.0 * 512;
      double fl_fma = (mode == 1 ? nw : mode == 2 ? nw / 2 : 0) * iters * 128.0 * 64;
      if (rep) printf("mode %d: %.3f ms  DMMA %.2f TF  DFMA %.2f TF  (sum %.2f)\n", mode, ms, fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
    }
  }
  return 0;
}
