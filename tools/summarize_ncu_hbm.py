#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` capture of the HBM-bound kernels
(tools/profile_hbm_kernels.py):

  ncu -i gpurun_out/hbm_kernels.ncu-rep --page raw --csv > gpurun_out/hbm_raw.csv
  python tools/summarize_ncu_hbm.py gpurun_out/hbm_raw.csv profiles/r1_ncu_full_hbm_kernels

writes <out>.txt (table) and <out>.json.  DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) /
gpu__time_duration.sum per launch; the peak is MEASURED_PEAKS.json's hbm_gbs.
"""
import collections
import csv
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}


def main(raw_csv, out):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = hdr.index

    def val(r, name):
        return float(r[col(name)]) * SCALE.get(units[col(name)], 1.0)

    peak = 6457.1
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        peak = float(json.load(open(ppath))["hbm_gbs"])
    agg = collections.OrderedDict()
    for r in data:
        key = (r[col("Kernel Name")].split("(")[0], r[col("launch__grid_size")], r[col("launch__block_size")])
        d = agg.setdefault(key, collections.defaultdict(list))
        d["t"].append(val(r, "gpu__time_duration.sum"))
        d["rd"].append(val(r, "dram__bytes_read.sum"))
        d["wr"].append(val(r, "dram__bytes_write.sum"))
        d["sm"].append(float(r[col("sm__throughput.avg.pct_of_peak_sustained_elapsed")]))
        d["occ"].append(float(r[col("sm__warps_active.avg.pct_of_peak_sustained_active")]))
        d["regs"].append(float(r[col("launch__registers_per_thread")]))
        d["ld_sec"].append(float(r[col("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")]))
        d["ld_req"].append(float(r[col("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum")]))
    lines = ["ncu --set full --clock-control none; per-launch means; DRAM GB/s = (dram__bytes_read.sum + "
             "dram__bytes_write.sum) / gpu__time_duration.sum; peak = MEASURED_PEAKS.json hbm_gbs %.1f" % peak,
             "%-28s %6s %5s %3s %8s %8s %8s %8s %7s %7s %7s %5s %11s" % (
                 "kernel", "grid", "block", "n", "time_us", "rd_MB", "wr_MB", "GB/s", "%peak", "sm_thr%", "warps%",
                 "regs", "sect/req_ld")]
    out_json = {}
    for (name, grid, blk), d in agg.items():
        t, rd, wr = np.mean(d["t"]), np.mean(d["rd"]), np.mean(d["wr"])
        gbs = (rd + wr) / t / 1e9
        spr = np.mean(d["ld_sec"]) / max(np.mean(d["ld_req"]), 1.0)
        lines.append("%-28s %6s %5s %3d %8.2f %8.3f %8.3f %8.1f %7.2f %7.1f %7.1f %5d %11.2f" % (
            name, grid, blk, len(d["t"]), t * 1e6, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak, np.mean(d["sm"]),
            np.mean(d["occ"]), np.mean(d["regs"]), spr))
        out_json["%s<<<%s,%s>>>" % (name, grid, blk)] = {
            "launches": len(d["t"]), "time_us": t * 1e6, "dram_read_bytes": rd, "dram_write_bytes": wr,
            "dram_GBps": gbs, "frac_of_hbm_peak": gbs / peak, "sm_throughput_pct": float(np.mean(d["sm"])),
            "warps_active_pct": float(np.mean(d["occ"])), "global_load_sectors_per_request": float(spr)}
    text = "\n".join(lines) + "\n"
    sys.stdout.write(text)
    with open(out + ".txt", "w") as f:
        f.write(text)
    with open(out + ".json", "w") as f:
        json.dump(out_json, f, indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
