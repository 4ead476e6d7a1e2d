"""k_trigemm_chi2: time and DRAM traffic against the super-group size (column tiles whose R panels are scheduled
together).  `python tools/r2_group_sweep.py time` prints the CUDA-event times; under
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_trigemm` the `ncu` mode
launches every variant twice (the second launch of each pair is the one to read)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import r2_chi2_sweep as S  # noqa: E402

GROUPS = [8, 16, 32, 48, 64, 96, 146, 256]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "time"
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    prog, prob = S.program("planck")
    for g in GROUPS:
        if mode == "ncu":
            S.time_chi2(prog, W, iters=1, pipe=1, trigemm_group=g)      # 3 warm-up launches + 1
            print(json.dumps(dict(W=W, group=g, launches=4)), flush=True)
        else:
            ms, _ = S.time_chi2(prog, W, iters=20, pipe=1, trigemm_group=g)
            print(json.dumps(dict(W=W, group=g, ms=round(ms, 5))), flush=True)


if __name__ == "__main__":
    main()
