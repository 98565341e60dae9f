#!/usr/bin/env python
"""profiles/<round>_traffic.json from an `ncu --set full` capture of one bench step taken at --spp 16 (one chunk):
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the closest-hit launches over the rays they traced.
  python tools/make_traffic.py capture.ncu-rep bench_line_of_the_same_command.json out.json"""
import csv, json, subprocess, sys
rep, line, out = sys.argv[1:4]
d = json.loads([l for l in open(line) if l.startswith("{")][-1])
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
units = rows[1]
ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram = [float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]] for r in rows[2:] if "k_closest" in r[ki]]
col = lambda name: [float(r[h.index(name)]) for r in rows[2:] if "k_closest" in r[ki]]
issue = col("smsp__issue_active.avg.pct_of_peak_sustained_active")
lanes = col("smsp__thread_inst_executed_per_inst_executed.ratio")
l1hit, l2hit = col("l1tex__t_sector_hit_rate.pct"), col("lts__t_sector_hit_rate.pct")
dur_ms = [float(r[h.index("gpu__time_duration.sum")]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[h.index("gpu__time_duration.sum")]] for r in rows[2:] if "k_closest" in r[ki]]
samples = d["samples_per_s"] * d["ms_per_step"] / 1e3
closest = d["closest_Mrays_per_s"] * 1e6 * d["ms_per_step"] / 1e3
rays = [samples, closest - samples][:len(dram)]       # one chunk: the camera-ray launch, then the bounce launch
json.dump({"kernel": "k_closest", "source": f"{rep} (ncu --set full, {d['config']['workload']})", "dram_bytes": dram, "rays": rays,
           "dram_bytes_per_ray": sum(dram) / sum(rays),
           "launches": ["camera rays", "bounce rays"][:len(dram)], "ncu_duration_ms": dur_ms, "issue_slots_busy_pct": issue, "active_lanes_per_instruction": lanes,
           "l1_hit_pct": l1hit, "l2_hit_pct": l2hit}, open(out, "w"), indent=1)
print(open(out).read())
