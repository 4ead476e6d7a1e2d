"""Summarise ncu outputs into profiles/ (run here, no GPU): the launch list (per-kernel share of the
step) and the raw-page metrics + SASS opcode mix + stall reasons of a --set full capture."""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn])[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = ["kernel,launches,total_us,share_pct,avg_us"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s,%d,%.1f,%.2f,%.1f" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot, v[1] / v[0] / 1e3))
    return "\n".join(out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum"]


def ncu_page(rep, page, kernel):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", "--kernel-name", "regex:" + kernel],
                          capture_output=True, text=True).stdout


def full(rep, kernel):
    rows = list(csv.reader(io.StringIO(ncu_page(rep, "raw", kernel))))
    h = rows[0]
    out = ["# %s (first captured launch)" % kernel]
    for w in WANT:
        if w in h:
            i = h.index(w)
            out.append("%s = %s %s" % (w, rows[2][i], rows[1][i]))
    rows = list(csv.reader(io.StringIO(ncu_page(rep, "source", kernel))))
    if len(rows) > 2:
        h = rows[1]
        ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        ops, samp = collections.Counter(), collections.Counter()
        for r in rows[2:]:
            if len(r) <= iex or not (r[iex] or "0").isdigit():
                continue
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += int(r[iex] or 0)
            samp[op] += int(r[isamp] or 0)
        te, ts = sum(ops.values()), max(sum(samp.values()), 1)
        out.append("## SASS opcode mix (warp instructions executed, stall samples)")
        for op, c in ops.most_common(12):
            out.append("%-8s exec %5.1f%%  samples %5.1f%%" % (op, 100.0 * c / te, 100.0 * samp[op] / ts))
        st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not" not in x]
        tot = collections.Counter()
        for r in rows[2:]:
            for i in st:
                if i < len(r) and r[i].isdigit():
                    tot[h[i]] += int(r[i])
        out.append("## stall reasons (samples): " + ", ".join("%s %d" % kv for kv in tot.most_common(8)))
    return "\n".join(out)


if __name__ == "__main__":
    what = sys.argv[1]
    if what == "launches":
        print(launches(sys.argv[2]))
    else:
        for k in sys.argv[3:]:
            print(full(sys.argv[2], k))
            print()
