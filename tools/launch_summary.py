#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total
device time and SHARE (absolute times under ncu are cold-cache and serialised).
usage: launch_summary.py launches.csv [note]"""
import collections, csv, sys

def main():
    path = sys.argv[1]; note = sys.argv[2] if len(sys.argv) > 2 else ""
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines); hdr = next(r)
    agg = collections.OrderedDict(); n = 0
    for row in r:
        if len(row) != len(hdr): continue
        d = dict(zip(hdr, row)); n += 1
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        a = agg.setdefault(d["Kernel Name"][:110], [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.1f} us total device time (under ncu: compare SHARES)\n# {note}")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:12.1f} us {100 * t / tot:6.2f}%  x{c:<5d} avg {t / c / 1e3:9.1f} us  {k}")

main()
