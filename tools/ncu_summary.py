"""Summaries of ncu exports. Usage: ncu_summary.py launches <launch-list csv> | raw <ncu-rep> (needs ncu on PATH)"""
import collections
import csv
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]


def launches(path):
    rows, hdr, agg = list(csv.reader(open(path))), None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            name = re.sub(r"\(.*", "", d["Kernel Name"])
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[:100]} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {a[1] / tot:.3f} |")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(d.get("Kernel Name", "")[:120])
        for w in WANT:
            if w in d:
                print(f"  {w}: {d[w]}")
        for k, v in d.items():
            if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                try:
                    if float(v) >= 0.2:
                        print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {float(v):.2f}")
                except ValueError:
                    pass


if __name__ == "__main__":
    (launches if sys.argv[1] == "launches" else raw)(sys.argv[2])
