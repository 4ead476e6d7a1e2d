#!/usr/bin/env python
"""Registers / stack frame / spills of every kernel, from `nvcc -Xptxas -v` (runs without a GPU):
    python tools/ptxas_resources.py > profiles/r1_ptxas_resources.txt
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "cosmoslik_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def main():
    print("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v; one line per kernel")
    with tempfile.TemporaryDirectory() as tmp:
        for f in ("prepare", "samplers", "chi2", "covchi2", "binning", "api"):
            r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                                "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-c", os.path.join(SRC, f + ".cu"), "-o",
                                os.path.join(tmp, f + ".o")], capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stderr)
                raise SystemExit(r.returncode)
            lines, name = r.stderr.splitlines(), None
            for i, l in enumerate(lines):
                m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", l)
                if m:
                    name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
                    name = re.sub(r"\(.*", "", name)
                m2 = re.search(r"Used (\d+) registers.*", l)
                if m2 and name:
                    st = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", lines[i - 1])
                    print("%-10s %-32s %-75s | stack %s B, spill st %s B, ld %s B" % (
                        (f, name[:32], l.split("ptxas info    : ")[-1][:75]) + (st.groups() if st else ("?", "?", "?"))))
                    name = None


if __name__ == "__main__":
    main()
