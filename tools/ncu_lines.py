#!/usr/bin/env python
"""Source-line view of one kernel of an ncu report (needs -lineinfo and --import-source on): lines ranked by executed warp
instructions and by stall samples.   tools/ncu_lines.py rep.ncu-rep <launch index> [n]"""
import csv, subprocess, sys, os
rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, h, lines = "?", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = os.path.basename(r[1]); continue
    if len(r) == 2 and r[0] == "Function Name":
        fn = r[1][:80]; continue
    if r and r[0] == "Line No":
        h = {n: i for i, n in enumerate(r)}; continue
    if h and len(r) > 8 and r[0].strip().isdigit():
        try:
            lines.append((cur, int(r[0]), r[1].strip(), int(r[h["Instructions Executed"]]), int(r[h["# Samples"]]), int(r[h["Thread Instructions Executed"]])))
        except ValueError:
            pass
ti = sum(l[3] for l in lines) or 1; ts = sum(l[4] for l in lines) or 1
print(fn, "| warp instr", ti, "samples", ts)
key = (lambda l: -l[4]) if len(sys.argv) > 4 and sys.argv[4] == "samples" else (lambda l: -l[3])
for f, n, src, i, s_, t in sorted(lines, key=key)[:top]:
    print("%-18s %5d inst %5.2f%% samp %5.2f%% lanes %4.1f | %s" % (f[:18], n, 100.0 * i / ti, 100.0 * s_ / ts, t / max(i, 1), src[:110]))
