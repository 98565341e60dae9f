"""Summarise ncu outputs for profiles/: (1) a launch-list CSV (gpu__time_duration per launch) -> per-kernel totals and
shares, (2) a --set full .ncu-rep -> key metrics per captured launch.   python tools/ncu_summary.py launches.csv prof.ncu-rep"""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        n = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    out = ["| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot:.3f} |")
    return "\n".join(out)


WANT = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed.sum.per_cycle_active", "warp inst / cycle / chip"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch resolving / issue"),
]


def report(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    names = [r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "") for r in data]
    out = ["| metric | " + " | ".join(f"{i}: `{n}`" for i, n in enumerate(names)) + " |", "|---|" + "---:|" * len(names)]
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            out.append(f"| {label} [{units[i]}] | " + " | ".join(r[i] for r in data) + " |")
    return "\n".join(out)


if __name__ == "__main__":
    for a in sys.argv[1:]:
        print(f"\n### {a}\n")
        print(launches(a) if a.endswith(".csv") else report(a))
