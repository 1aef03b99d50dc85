#!/usr/bin/env python3
"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
Usage: summarize_launches.py launches.csv [title] > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r.get("Metric Unit") in ("us", "usecond"):
            ns *= 1e3
        elif r.get("Metric Unit") in ("ms", "msecond"):
            ns *= 1e6
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("nq::", "").replace("void ", "").replace("<unnamed>::", "")
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    agg = collections.OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, [0, 0.0, grid, block])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"# {title}\n")
    print(f"{len(rows)} launches, {total / 1e6:.3f} ms of kernel time (ncu per-launch times are serialised and cold-cache: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total ms | share | grid | block |")
    print("|---|---:|---:|---:|---|---|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / total:.2f}% | {a[2]} | {a[3]} |")


if __name__ == "__main__":
    main()
