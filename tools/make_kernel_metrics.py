#!/usr/bin/env python
"""profiles/<round>_kernel_metrics.json from an `ncu --set full` capture of ONE round of a bench.py command and the JSON line
of the same command run without ncu: per kernel class (closest / shadow / shade / sampler) DRAM bytes and warp instructions per
unit (ray, k_shade thread, pixel), active lanes per instruction, occupancy, cache hit rates -- what bench.py turns into
frac_hbm and frac_issue with its live timings -- plus one row per captured launch (the ncu inventory of the round).
  python tools/make_kernel_metrics.py capture.ncu-rep plain_line.json out.json [summary.md]"""
import csv, json, subprocess, sys
rep, line, out = sys.argv[1:4]
d = json.loads([l for l in open(line) if l.startswith("{")][-1])
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h, units = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "": 1, "inst": 1, "%": 1}


def val(r, name):
    i = col[name]
    try:
        return float(r[i].replace(",", "")) * scale.get(units[i], 1)
    except ValueError:
        return 0.0


def cls_of(name):
    if "k_closest" in name: return "closest"
    if "k_shadow" in name: return "shadow"
    if "k_sampler" in name or "k_pixel_setup" in name: return "sampler"
    if "k_shade" in name or "k_raygen" in name or "k_bin" in name or "k_finish" in name: return "shade"
    return None


launches, agg = [], {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    rec = {"kernel": short, "ms": val(r, "gpu__time_duration.sum"), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
           "registers": val(r, "launch__registers_per_thread"),
           "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
           "dram_pct_of_peak": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
           "l2_pct_of_peak": val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
           "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct"),
           "warp_inst": val(r, "smsp__inst_executed.sum"), "lanes_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
           "issue_slots_busy_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "achieved_occupancy_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")}
    rec["dram_GBs"] = rec["dram_bytes"] / (rec["ms"] / 1e3) / 1e9 if rec["ms"] else 0.0
    for name2 in ("lts__t_bytes.sum", "lts__t_sectors.sum"):
        if name2 in col:
            b = val(r, name2) * (32 if name2.endswith("sectors.sum") else 1)
            rec["l2_GBs"] = b / (rec["ms"] / 1e3) / 1e9 if rec["ms"] else 0.0
            break
    launches.append(rec)
    c = cls_of(name)
    if c:
        a = agg.setdefault(c, {"ms": 0.0, "dram_bytes": 0.0, "warp_inst": 0.0, "thread_inst": 0.0, "regs": 0, "occ_w": 0.0, "l1_w": 0.0, "l2_w": 0.0, "kernels": []})
        a["ms"] += rec["ms"]; a["dram_bytes"] += rec["dram_bytes"]; a["warp_inst"] += rec["warp_inst"]
        a["thread_inst"] += rec["warp_inst"] * rec["lanes_per_inst"]
        a["regs"] = max(a["regs"], int(rec["registers"]))
        a["occ_w"] += rec["achieved_occupancy_pct"] * rec["ms"]; a["l1_w"] += rec["l1_hit_pct"] * rec["ms"]; a["l2_w"] += rec["l2_hit_pct"] * rec["ms"]
        if short not in a["kernels"]: a["kernels"].append(short)
# units of the captured round, from the plain run of the same command (one timed step)
steps = max(1, d["steps"])
samples = d["samples_per_s"] * d["ms_per_step"] / 1e3
rays_total = d["value"] * 1e6 * d["ms_per_step"] / 1e3
closest = d.get("closest_rays_per_step") or 0
shadow = rays_total - closest
pixels = samples / d["spp"]
unit_count = {"closest": closest, "shadow": shadow, "shade": closest, "sampler": pixels}
res = {"source": "%s (ncu --set full --clock-control none, one round of: %s)" % (rep.split("/")[-1], d["config"]["workload"]),
       "units_of_the_captured_round": unit_count}
for c, a in agg.items():
    u = max(1.0, unit_count[c])
    res[c] = {"kernels": a["kernels"], "ncu_ms": a["ms"], "dram_bytes_per_unit": a["dram_bytes"] / u, "warp_inst_per_unit": a["warp_inst"] / u,
              "active_lanes_per_instruction": a["thread_inst"] / max(1.0, a["warp_inst"]), "registers_per_thread": a["regs"],
              "achieved_occupancy_pct": a["occ_w"] / max(1e-9, a["ms"]), "l1_hit_pct": a["l1_w"] / max(1e-9, a["ms"]), "l2_hit_pct": a["l2_w"] / max(1e-9, a["ms"])}
res["launches"] = launches
json.dump(res, open(out, "w"), indent=1)
if len(sys.argv) > 4:
    with open(sys.argv[4], "w") as f:
        f.write("| kernel | grid x block | regs | ms | DRAM GB/s (%% of peak) | L2 GB/s (%% of peak) | L1 / L2 hit %% | warp inst | lanes / inst | issue slots busy %% | occupancy %% |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in launches:
            f.write("| `%s` | %s x %s | %d | %.3f | %.0f (%.1f) | %.0f (%.1f) | %.0f / %.0f | %.3g | %.1f | %.1f | %.1f |\n" % (
                r["kernel"], r["grid"], r["block"], r["registers"], r["ms"], r["dram_GBs"], r["dram_pct_of_peak"], r.get("l2_GBs", 0.0), r["l2_pct_of_peak"],
                r["l1_hit_pct"], r["l2_hit_pct"], r["warp_inst"], r["lanes_per_inst"], r["issue_slots_busy_pct"], r["achieved_occupancy_pct"]))
print(json.dumps({k: v for k, v in res.items() if k != "launches"}, indent=1))
