"""profiles/r2_traffic.json from this round's `ncu --set full` summaries (profiles/r2_ncu_full_*.txt, produced by
tools/r2_profile.sh + tools/ncu_summary.py): DRAM bytes per launch of the two FP64 kernels of the headline, next to
the algorithmic bytes, stamped with the digest of the kernel sources the captures were taken from.  bench.py quotes
`roofline.traffic` from this file only while the digest and the launch size still match.  No GPU needed."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def dram_bytes(path):
    tot = 0.0
    for ln in open(path):
        m = re.match(r"dram__bytes_(read|write)\.sum = ([0-9.]+) (\w+)", ln)
        if m:
            tot += float(m.group(2)) * UNIT[m.group(3)]
    return int(tot)


def main():
    walkers = 32768
    n, npad = 613, 640
    out = {"source_digest": bench.kernel_source_digest(),
           "k_trigemm_chi2": {
               "walkers_per_launch": walkers,
               "dram_bytes_per_launch": dram_bytes(os.path.join(ROOT, "profiles", "r2_ncu_full_k_trigemm_chi2.txt")),
               "algorithmic_bytes_per_launch": n * walkers * 8 + npad * npad * 8,
               "source": "ncu --set full, profiles/r2_ncu_full_k_trigemm_chi2.txt (dram__bytes_read.sum + "
                         "dram__bytes_write.sum of the first captured launch); algorithmic = R read once "
                         "(613 x 32768 x 8 B) + L^-1 (640^2 x 8 B)"},
           "k_bin_segsum": {
               "walkers_per_launch": walkers,
               "dram_bytes_per_launch": dram_bytes(os.path.join(ROOT, "profiles", "r2_ncu_full_k_bin_segsum.txt")),
               "algorithmic_bytes_per_launch": npad * walkers * 8,
               "source": "profiles/r2_ncu_full_k_bin_segsum.txt (k_bin_segsum_pair: both classes of a half-step in one "
                         "launch; algorithmic = R written once, 640 x 32768 x 8 B -- part of it is still in L2 when "
                         "the launch ends, so the capture counts fewer DRAM bytes than that)"}}
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
