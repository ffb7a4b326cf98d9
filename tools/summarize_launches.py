"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.

    python tools/summarize_launches.py gpurun_out/launches.csv [--steps N] > profiles/launches_summary.md
"""
import argparse
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<.*$", "", name)
    return name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--steps", type=int, default=1, help="training steps covered by the capture")
    a = ap.parse_args()
    rows = []
    with open(a.csv, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    tot = defaultdict(lambda: [0.0, 0])
    for n, t in rows:
        k = short(n)
        tot[k][0] += t
        tot[k][1] += 1
    total = sum(v[0] for v in tot.values())
    print(f"launches: {len(rows)}  total kernel time: {total / 1e6:.3f} ms  ({total / 1e6 / a.steps:.3f} ms per step over {a.steps} steps)\n")
    print("| kernel | launches/step | us/launch | ms/step | share |")
    print("|---|---:|---:|---:|---:|")
    for k, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
        ours = "**" if k.startswith("pcb::") else ""
        print(f"| {ours}{k}{ours} | {c / a.steps:.1f} | {t / c / 1e3:.2f} | {t / 1e6 / a.steps:.3f} | {t / total:.3f} |")
    mine = sum(v[0] for k, v in tot.items() if k.startswith("pcb::"))
    print(f"\nkernels of libpcbridge (pcb::*): {mine / total:.3f} of kernel time")


if __name__ == "__main__":
    main()
