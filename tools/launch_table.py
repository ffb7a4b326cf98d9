"""Per-launch table of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` capture:

    python tools/launch_table.py gpurun_out/step.csv [substring ...]      # only kernels whose name contains a substring
"""
import csv
import re
import sys
from collections import defaultdict


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    L = {}
    for r in csv.DictReader(lines):
        d = L.setdefault(int(r["ID"]), {"k": re.sub(r"\(.*$", "", re.sub(r"^void ", "", r["Kernel Name"]))[:56],
                                        "grid": r["Grid Size"], "blk": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return L


if __name__ == "__main__":
    L = load(sys.argv[1])
    subs = sys.argv[2:]
    tot = defaultdict(lambda: [0.0, 0])
    for i, d in sorted(L.items()):
        t = d.get("gpu__time_duration.sum", 0.0) / 1e3
        tot[d["k"]][0] += t
        tot[d["k"]][1] += 1
        if subs and any(s in d["k"] for s in subs):
            print("%4d %-56s %-16s %6.1f us  rd %6.1f MB  wr %6.1f MB" % (
                i, d["k"], d["grid"], t, d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6))
    T = sum(v[0] for v in tot.values())
    print("launches %d, kernel time %.1f us" % (sum(v[1] for v in tot.values()), T))
    for k, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:30]:
        print("  %-58s n=%3d  us/launch=%7.2f  total=%8.1f  share=%.3f" % (k, c, t / c, t, t / T))
