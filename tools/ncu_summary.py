"""Markdown table of the key metrics of every kernel in an `ncu --set full` report:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import re
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "tcgen05 issue %"),
        ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM ld/st issue %"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMA unit active %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %")]


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {n: hdr.index(n) for n, _ in COLS if n in hdr}
    kn = hdr.index("Kernel Name")
    print("| kernel | " + " | ".join(f"{t} [{units[idx[n]]}]" if units[idx[n]] else t for n, t in COLS if n in idx) + " |")
    print("|---|" + "---:|" * len(idx))
    for d in data:
        name = re.sub(r"^void (pcb::)?", "", d[kn])
        name = re.sub(r"\(.*", "", name)[:60]
        vals = []
        for n, _ in COLS:
            if n in idx:
                v = d[idx[n]]
                try:
                    v = f"{float(v.replace(',', '')):.4g}"
                except ValueError:
                    pass
                vals.append(v)
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
