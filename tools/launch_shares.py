#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, mean, share.

    python tools/launch_shares.py gpurun_out/launches.csv [-o profiles/name.txt]
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    out = sys.argv[sys.argv.index("-o") + 1] if "-o" in sys.argv else None
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, rows = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0][-90:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["%s: %d launches, %.1f us of kernel time (ncu per-launch times are cold-cache and serialised: read the shares)" % (path, len(rows), tot)]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-92s n=%5d total=%10.1f us mean=%9.2f us share=%5.1f%%" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
    text = "\n".join(lines) + "\n"
    if out:
        open(out, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
