"""ncu per-launch metrics of ONE eager train step (tools/ncu_step.py) -> per-kernel table (markdown, stdout) and
profiles/traffic_per_launch.json (average DRAM bytes per launch of each libpcbridge kernel; bench.py reports it as
`roofline.traffic`).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \\
        --clock-control none --profile-from-start off --csv --log-file gpurun_out/step.csv python tools/ncu_step.py
    python tools/step_traffic.py gpurun_out/step.csv profiles/traffic_per_launch.json > profiles/step_kernels.md
"""
import csv
import json
import re
import sys
from collections import defaultdict

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"<.*$", "", name)
    return name


def main():
    src, out_json = sys.argv[1], sys.argv[2]
    lines = [l for l in open(src, newline="") if not l.startswith("==")]
    per = defaultdict(lambda: defaultdict(float))
    count = defaultdict(int)
    seen = set()
    for r in csv.DictReader(lines):
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
        per[k][r["Metric Name"]] += v
        if (r["ID"], k) not in seen:
            seen.add((r["ID"], k))
            count[k] += 1
    total_us = sum(m["gpu__time_duration.sum"] for m in per.values())
    print("| kernel | launches | us/launch | share of kernel time | DRAM read MB/launch | DRAM write MB/launch | L2 hit % |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    traffic = {}
    for k, m in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = count[k]
        rd, wr = m["dram__bytes_read.sum"] / n, m["dram__bytes_write.sum"] / n
        ours = k.startswith("pcb::")
        if ours:
            traffic[k[5:]] = int(rd + wr)
        print(f"| {'**' if ours else ''}`{k[:64]}`{'**' if ours else ''} | {n} | {m['gpu__time_duration.sum'] / n:.1f} | "
              f"{m['gpu__time_duration.sum'] / total_us:.3f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | {m['lts__t_sector_hit_rate.pct'] / n:.0f} |")
    print(f"\nall launches: {sum(count.values())}, kernel time {total_us / 1e3:.3f} ms")
    traffic["_source"] = ("ncu (dram__bytes_read.sum + dram__bytes_write.sum, --clock-control none) over every launch of ONE eager "
                          "MSG train step (tools/ncu_step.py), averaged per kernel; " + src)
    json.dump(traffic, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
