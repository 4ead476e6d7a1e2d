"""Measure the FP64 roofline denominators on the B200 (run under gpurun; writes gpurun_out/):
cuBLAS DGEMM through torch.matmul (burst = best of 10; sustained = back to back for ~3 s) and the
package's own DMMA tile kernel (csl_dgemm_dmma) for comparison."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cosmoslik_b200 import _lib as L


def time_fn(fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    torch.cuda.set_device(0)
    N = 8192
    A = torch.randn(N, N, dtype=torch.float64, device="cuda")
    B = torch.randn(N, N, dtype=torch.float64, device="cuda")
    Cm = torch.empty(N, N, dtype=torch.float64, device="cuda")
    fl = 2.0 * N ** 3
    for _ in range(3):
        torch.matmul(A, B, out=Cm)
    burst = min(time_fn(lambda: torch.matmul(A, B, out=Cm), 1) for _ in range(10))
    t0 = time.time(); reps = 0; tot = 0.0
    while time.time() - t0 < 3.0:
        tot += time_fn(lambda: torch.matmul(A, B, out=Cm), 4) * 4; reps += 4
    out = {"fp64_tflops_burst": fl / burst / 1e9, "fp64_tflops_sustained": fl / (tot / reps) / 1e9,
           "how": "torch.matmul float64 8192^3 (cuBLAS DGEMM): best of 10 and back to back for 3 s",
           "gpu": torch.cuda.get_device_name(0)}
    lib = L.lib()
    M, Nn, K = 4096, 16384, 4096
    A2 = torch.randn(M, K, dtype=torch.float64, device="cuda")
    B2 = torch.randn(K, Nn, dtype=torch.float64, device="cuda")
    C2 = torch.empty(M, Nn, dtype=torch.float64, device="cuda")
    call = lambda: L.check(lib.csl_dgemm_dmma(M, Nn, K, A2.data_ptr(), B2.data_ptr(), C2.data_ptr(), L.stream_ptr()))
    for _ in range(3):
        call()
    ms = min(time_fn(call, 2) for _ in range(5))
    out["own_dmma_tile_gemm_tflops"] = 2.0 * M * Nn * K / ms / 1e9
    err = (C2[:64, :64] - (A2[:64] @ B2[:, :64])).abs().max().item()
    out["own_dmma_max_abs_err"] = err
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/measured_fp64_peak.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
