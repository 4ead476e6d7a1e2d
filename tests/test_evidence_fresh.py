"""The committed ncu evidence must belong to the committed kernels: profiles/r2_traffic.json is stamped with a digest of
cosmoslik_b200/csrc/*.cu(h); bench.py quotes `roofline.traffic` from it only while the digest matches.  A kernel edit
without a fresh capture (tools/r2_profile.sh + tools/r2_traffic.py) fails here instead of silently nulling the field."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_traffic_capture_matches_the_kernel_sources():
    import bench
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
        tj = json.load(f)
    assert tj["source_digest"] == bench.kernel_source_digest()
    rec = tj["k_trigemm_chi2"]
    assert rec["walkers_per_launch"] == bench.LITERAL_WALKERS // 2
    # the dominant kernel reads its operands about once: traffic within 10 % of the algorithmic bytes
    assert rec["algorithmic_bytes_per_launch"] <= rec["dram_bytes_per_launch"] <= 1.10 * rec["algorithmic_bytes_per_launch"]
    assert bench.measured_traffic("k_trigemm_chi2", bench.LITERAL_WALKERS // 2) == rec["dram_bytes_per_launch"]


def test_both_bench_arms_describe_the_same_config():
    import types
    import bench
    args = types.SimpleNamespace(walkers=bench.LITERAL_WALKERS, gpus=1, workload="planck_lite", binning="auto")
    prob = {"n": 613, "truth": dict.fromkeys("abcdefghi")}
    a, b = bench.config_of(args, prob), bench.config_of(args, prob)
    assert a == b and a["walkers_total"] == 65536 and "model" not in a
