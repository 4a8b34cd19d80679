#!/usr/bin/env python
"""Turns an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list into a per-kernel
markdown table (profiles/*.md).  Usage: summarize_launches.py launches.csv "command that was profiled" > out.md"""
import collections
import csv
import re
import sys


def main():
    path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    t = collections.defaultdict(float)
    n = collections.Counter()
    dram = collections.defaultdict(float)
    for r in csv.DictReader(lines[start:]):
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "")
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
            t[k] += v
            n[k] += 1
        elif r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
            dram[k] += v
    tot = sum(t.values())
    print(f"# ncu launch list: `{cmd}`\n")
    print("Per-launch `gpu__time_duration.sum` (cold-cache, serialised under ncu: shares matter, not absolutes).\n")
    print(f"Total kernel time {tot / 1e3:.2f} ms over {sum(n.values())} launches.\n")
    print("| kernel | launches | total ms | share | avg us | DRAM GB |")
    print("|---|---:|---:|---:|---:|---:|")
    for k, v in sorted(t.items(), key=lambda x: -x[1]):
        print(f"| `{k}` | {n[k]} | {v / 1e3:.3f} | {100 * v / tot:.1f}% | {v / n[k]:.1f} | {dram[k] / 1e9:.2f} |")


if __name__ == "__main__":
    main()
