#!/usr/bin/env python
"""Per-instruction view of one kernel of an ncu report: tools/ncu_sass.py rep.ncu-rep <launch index> [min_share]
Prints SASS lines with executed warp instructions, average active threads and stall samples (source page)."""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = {n: i for i, n in enumerate(rows[hi])}
print(rows[0][:2])
body = [r for r in rows[hi + 1:] if len(r) > 5 and r[h["Instructions Executed"]].isdigit()]
tot_i = sum(int(r[h["Instructions Executed"]]) for r in body)
tot_s = sum(int(r[h["# Samples"]]) for r in body)
print("total warp instr", tot_i, "samples", tot_s)
for r in body:
    ie = int(r[h["Instructions Executed"]]); sm = int(r[h["# Samples"]])
    print("%s %-62s inst %5.2f%% thr %5s samp %5.2f%% lsb %s" % (r[0][-4:], r[1].strip()[:62], 100.0 * ie / tot_i, r[h["Avg. Threads Executed"]][:5], 100.0 * sm / max(tot_s, 1), r[h["stall_long_sb"]]))
