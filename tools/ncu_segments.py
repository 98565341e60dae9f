#!/usr/bin/env python
"""Basic-block view of one kernel of an ncu report: runs of SASS instructions with (nearly) equal execution counts, largest first,
with their instruction mix, share of executed warp instructions and of stall samples.
   tools/ncu_segments.py rep.ncu-rep <launch index> [n]"""
import csv, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = {n: i for i, n in enumerate(rows[hi])}
print(rows[0][1][:100])
body = [r for r in rows[hi + 1:] if len(r) > 5 and r[h["Instructions Executed"]].isdigit()]
ie, sm, th = h["Instructions Executed"], h["# Samples"], h["Thread Instructions Executed"]
tot = sum(int(r[ie]) for r in body); ts = sum(int(r[sm]) for r in body); tt = sum(int(r[th]) for r in body)
print("SASS instructions", len(body), "executed warp instr", tot, "avg lanes %.1f" % (tt / max(tot, 1)), "samples", ts)
segs, i = [], 0
while i < len(body):
    j, c = i, int(body[i][ie])
    while j < len(body) and abs(int(body[j][ie]) - c) <= 0.02 * max(c, 1):
        j += 1
    segs.append((i, j, c, sum(int(r[ie]) for r in body[i:j]), sum(int(r[sm]) for r in body[i:j]), sum(int(r[th]) for r in body[i:j])))
    i = j
for i, j, c, ti, tsm, tth in sorted(segs, key=lambda s: -max(s[3] / tot, s[4] / max(ts, 1)))[:top]:
    ops = {}
    for r in body[i:j]:
        w = r[1].split()
        op = (w[1] if w[0].startswith("@") else w[0]).split(".")[0]
        ops[op] = ops.get(op, 0) + 1
    mix = " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:7])
    print(f"[{i}:{j}] n={j - i} exec={c / 1e6:.2f}M lanes={tth / max(ti, 1):.1f} inst {100 * ti / tot:.1f}% samp {100 * tsm / max(ts, 1):.1f}% | {mix}")
