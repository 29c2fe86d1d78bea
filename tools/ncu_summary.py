#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.

    python tools/ncu_summary.py gpurun_out/launches.csv [--top 40] > profiles/rNN_launches_summary.md
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    rows = []
    with open(path, newline="") as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu, ig, ib = (hdr.index(n) for n in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
    for r in rd:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
        rows.append((name, ns, r[ig], r[ib]))
    total = sum(r[1] for r in rows)
    agg = OrderedDict()
    for name, ns, g, b in rows:
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += ns
        a[2] = max(a[2], ns)
    print("# ncu launch list summary: %s" % path)
    print("%d launches, %.3f ms of kernel time (cold-cache, serialised: compare SHARES)\n" % (len(rows), total / 1e6))
    print("| kernel | launches | total ms | share | avg us | max us |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, (n, ns, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f |" % (name, n, ns / 1e6, 100 * ns / total, ns / n / 1e3, mx / 1e3))


if __name__ == "__main__":
    main()
