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
samples = d["samples_per_s"] * d["ms_per_step"] / 1e3
closest = d["closest_Mrays_per_s"] * 1e6 * d["ms_per_step"] / 1e3
rays = [samples, closest - samples][:len(dram)]       # one chunk: the camera-ray launch, then the bounce launch
json.dump({"kernel": "k_closest", "source": f"{rep} (ncu --set full, {d['config']['workload']})", "dram_bytes": dram, "rays": rays,
           "dram_bytes_per_ray": sum(dram) / sum(rays)}, open(out, "w"), indent=1)
print(open(out).read())
