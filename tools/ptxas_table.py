#!/usr/bin/env python
"""Register / stack / spill / static-shared-memory table of every kernel in aldi_b200/csrc (CPU only: nvcc cross-compiles).

    python tools/ptxas_table.py > profiles/rNN_ptxas_v.txt
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "aldi_b200", "csrc")


def one(src):
    r = subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                        "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", "/dev/null"], capture_output=True, text=True)
    rows, name = [], None
    for ln in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", ln)
        if m:
            name = subprocess.run(["/usr/local/cuda/bin/cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "").replace("(anonymous namespace)::", "")
            stack = spill = 0
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", ln)
        if m:
            stack, spill = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", ln)
        if m and name:
            sm = re.search(r"(\d+) bytes smem", ln)
            rows.append((src, int(m.group(1)), stack, spill, int(sm.group(1)) if sm else 0, name))
            name = None
    return rows


def main():
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    with ThreadPoolExecutor(8) as ex:
        rows = [r for rs in ex.map(one, srcs) for r in rs]
    print("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v over aldi_b200/csrc/*.cu (static smem only;")
    print("# dynamic shared memory is set per launch).  kernels: %d, with register spills: %d" % (len(rows), sum(1 for r in rows if r[3])))
    print("# (the product build adds -fmad=false for select*.cu; register counts there may differ by a few)")
    print("%-15s %5s %6s %6s %8s  %s" % ("file", "regs", "stack", "spill", "smem(B)", "kernel"))
    for r in rows:
        print("%-15s %5d %6d %6d %8d  %s" % r)


if __name__ == "__main__":
    sys.exit(main())
