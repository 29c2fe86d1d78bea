#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
per-kernel launches, total time, share and (when captured) DRAM traffic.

    python tools/ncu_summary.py gpurun_out/launches.csv [--top 40] [--json out.json] > profiles/rNN_launches.md
"""
import csv
import json
import re
import sys
from collections import OrderedDict

UNIT_NS = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}
UNIT_B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    jout = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ii, ik, im, iv, iu = (hdr.index(n) for n in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    launches = OrderedDict()   # id -> {"name", "ns", "rd", "wr"}
    for r in rd:
        if len(r) <= iv:
            continue
        d = launches.setdefault(r[ii], {"name": re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", ""),
                                        "ns": 0.0, "rd": None, "wr": None})
        v = float(r[iv].replace(",", ""))
        if r[im] == "gpu__time_duration.sum":
            d["ns"] = v * UNIT_NS.get(r[iu], 1.0)
        elif r[im] == "dram__bytes_read.sum":
            d["rd"] = v * UNIT_B.get(r[iu], 1.0)
        elif r[im] == "dram__bytes_write.sum":
            d["wr"] = v * UNIT_B.get(r[iu], 1.0)
    rows = list(launches.values())
    total = sum(r["ns"] for r in rows)
    have_dram = any(r["rd"] is not None for r in rows)
    agg = OrderedDict()
    for r in rows:
        a = agg.setdefault(r["name"], {"n": 0, "ns": 0.0, "max": 0.0, "bytes": 0.0})
        a["n"] += 1
        a["ns"] += r["ns"]
        a["max"] = max(a["max"], r["ns"])
        a["bytes"] += (r["rd"] or 0.0) + (r["wr"] or 0.0)
    print("# ncu launch list summary: %s" % path)
    print("%d launches, %.3f ms of kernel time (cold-cache, serialised: compare SHARES)\n" % (len(rows), total / 1e6))
    if have_dram:
        print("| kernel | launches | total ms | share | avg us | max us | DRAM MB (rd+wr) | MB/launch | GB/s |")
        print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
    else:
        print("| kernel | launches | total ms | share | avg us | max us |")
        print("|---|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"])[:top]:
        base = "| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f |" % (name[:70], a["n"], a["ns"] / 1e6, 100 * a["ns"] / total,
                                                             a["ns"] / a["n"] / 1e3, a["max"] / 1e3)
        if have_dram:
            base += " %.1f | %.2f | %.0f |" % (a["bytes"] / 1e6, a["bytes"] / a["n"] / 1e6, a["bytes"] / max(a["ns"], 1.0))
        print(base)
    if have_dram:
        print("\ntotal DRAM traffic of the step: %.2f GB" % (sum(a["bytes"] for a in agg.values()) / 1e9))
    if jout:
        out = {}
        for name, a in agg.items():
            key = "conv_tc" if name.startswith("conv_tc_kernel") else "wgrad_tc" if name.startswith("wgrad_tc_kernel") else name[:60]
            o = out.setdefault(key, {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
            o["launches"] += a["n"]
            o["ms"] += a["ns"] / 1e6
            o["dram_bytes"] += a["bytes"]
        for o in out.values():
            o["dram_bytes_per_launch"] = o["dram_bytes"] / max(o["launches"], 1)
        json.dump({"source": path, "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                   "--clock-control none, one eager step (bench.py --ncu-step)", "kernels": out}, open(jout, "w"), indent=1)


if __name__ == "__main__":
    main()
