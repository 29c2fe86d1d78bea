#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report, one markdown table (reads the .ncu-rep with the local ncu).

    python tools/ncu_rep_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_ncu.md
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index("Kernel Name")
    print("# ncu --set full: %s\n" % rep)
    print("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |")
    print("|---|---|" + "---:|" * len(data))
    print("| kernel | | " + " | ".join("`%s`" % r[ik].split("(")[0].replace("void <unnamed>::", "") for r in data) + " |")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("| %s | %s | %s |" % (w, units[i], " | ".join(r[i] for r in data)))
    if "dram__bytes_read.sum" in hdr:
        ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        print("| traffic = read + write | %s | %s |" % (units[ir], " | ".join("%.3f" % (float(r[ir]) + float(r[iw])) for r in data)))


if __name__ == "__main__":
    main()
