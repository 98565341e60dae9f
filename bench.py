#!/usr/bin/env python
"""bench.py -- headline benchmark of the RGKrt hot path on B200.

Metric (BASELINE.json): Mrays/s, closest-hit + shadow rays, whole job.  Workload: BASELINE configs[1],
scenes/sponza.json at 1920x1080, 64 spp, path tracing with next-event estimation (recursion-max 2) -- on the
seeded ~71 k-triangle atrium STAND-IN, because sponza.obj is not distributed with the reference
(rgk_b200/standin.py).  One "step" = one RenderDriver round (every pixel, 64 spp) per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload sponza|cornell|sibenik|conference|dragon-sponza]
                  [--sampler mt|fast] [--shard rounds|tiles] [--traversal bvh|kd] [--cfg field=value,...] [--quick]

What runs is the library default: the 4-wide BVH candidate pass with the reference's kd-tree as arbiter (rgk_device_cfg,
no environment variable).  `value`: device-resident framebuffer, CUDA events, max over ranks.  `e2e`: the same step
through rgk_render_round with pinned HOST framebuffers, H2D + D2H inside the timed region.
N > 1 (torchrun): rounds are sharded (GPU g renders round step*N+g with its own seed base: weak scaling); each GPU renders
into a per-round partial buffer, ONE NCCL reduce per round runs on a second stream under the next round, and rank 0 adds
the result to the running total (EXRTexture::Accumulate on the device).  Sample counts are analytic and not reduced.  The
same run then times the tile-sharded (strong-scaling) variant and checks its image bit for bit against a single-GPU render
(`strong`).  --shard tiles makes the tile-sharded variant the headline instead.
Untimed legs of the N = 1 line: `parity` (the headline round on the kd-only traversal, framebuffers compared bit for bit),
a counting round for the algorithmic bytes (`roofline`, `roofline_by_kernel`) and the reference's CPU renderer on a strided
sample of the frame's own tiles (`cpu_baseline`).  `--impl reference` times that CPU renderer alone (oracle/_ref, else
the oracle port).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from rgk_b200 import abi, scenes, standin  # noqa: E402

METRIC = "Mrays/s (closest-hit+shadow)"
CPU_TILES = 128     # CPU baseline sample: this many 32x32 tiles taken at a constant stride through the frame's own task list, full spp
KERNEL_METRICS = os.path.join(ROOT, "profiles", "r2_kernel_metrics.json")   # per-class ncu figures (tools/make_kernel_metrics.py)


def real_asset_workload(name, spp):
    """BASELINE configs on the real assets when $RGK_ASSETS points at a copy of the reference's scenes/ directory that
    has the meshes (sponza.obj, sibenik.obj, ... are not distributed with the reference).  Returns None otherwise."""
    root = os.environ.get("RGK_ASSETS")
    sizes = {"sponza": (1920, 1080, 64), "sibenik": (1920, 1080, 256), "conference": (3840, 2160, 1024), "dragon-sponza": (3840, 2160, 512)}
    if not root or name not in sizes:
        return None
    cfg_path = os.path.join(root, name + ".json")
    if not os.path.exists(cfg_path):
        return None
    from rgk_b200 import assets, scene
    w, h, ms = sizes[name]
    try:
        pack, cfg = scene.load_json_config(cfg_path, overrides={"output-width": w, "output-height": h, "multisample": spp or ms, "rounds": 1},
                                           mesh_loader=assets.load_obj_into, texture_loader=assets.load_image)
    except (scene.ConfigFileException, OSError) as e:
        print("RGK_ASSETS: %s (%s); using the stand-in" % (cfg_path, e), file=sys.stderr)
        return None
    return pack, cfg, "scenes/%s.json %dx%d %dspp on the real asset (%d tris)" % (name, w, h, spp or ms, pack.n_triangles)


def build_workload(name, spp=None, res=None):
    real = real_asset_workload(name, spp)
    if real is not None:
        return real
    if name == "sponza":
        kw = {"multisample": spp} if spp else {}
        if res:
            kw["width"], kw["height"] = (int(x) for x in res.split("x"))
        pack, cfg = standin.sponza(**kw)
        label = "scenes/sponza.json 1920x1080 64spp NEE recursion-max 2 (atrium stand-in, %d tris)" % pack.n_triangles
    elif name == "sibenik":
        pack, cfg = standin.sibenik(**({"multisample": spp} if spp else {}))
        label = "scenes/sibenik.json 1920x1080 256spp lens+envmap (stand-in, %d tris)" % pack.n_triangles
    elif name == "conference":
        pack, cfg = standin.conference(**({"multisample": spp} if spp else {}))
        label = "scenes/conference.json 3840x2160 1024spp 8 sphere lights recursion-max 4 (stand-in, %d tris)" % pack.n_triangles
    elif name == "dragon-sponza":
        pack, cfg = standin.dragon_sponza(**({"multisample": spp} if spp else {}))
        label = "scenes/dragon-sponza.json 3840x2160 512spp (stand-in, %d tris)" % pack.n_triangles
    elif name == "cornell":
        pack, cfg = scenes.load_builtin("cornell-box", **({"multisample": spp} if spp else {}))
        label = "scenes/cornell-box.json 256x256 16spp recursion-max 40"
    else:
        raise SystemExit("unknown workload " + name)
    if spp:
        label += " [spp overridden to %d]" % spp
    if res and name == "sponza":
        label += " [resolution overridden to %s]" % res
    return pack, cfg, label


def parse_cfg(text):
    out = {}
    for item in filter(None, (text or "").split(",")):
        k, v = item.split("=")
        out[k.strip()] = float(v) if "." in v else int(v)
    return out


class ClockSampler:
    """The recipe's clocks line (B200_PROFILING.md): ONE nvidia-smi process looping every 200 ms while the timed region runs --
    started before, killed after.  (Round 1 spawned a new nvidia-smi every 100 ms; each start-up takes the driver's global
    lock for ~0.2 s, which stalled this process's kernel launches and cost ~17 % of the headline on a 16-core box.)"""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.out = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=self.out, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.out.flush()
        self.out.seek(0)
        rows = [[x.strip() for x in line.split(",")] for line in self.out.read().splitlines() if line.strip()]
        self.out.close()
        try:
            os.unlink(self.out.name)
        except OSError:
            pass
        rows = [r for r in rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(rows[0][1]) if rows[0][1].isdigit() else None, "reasons": reasons,
                "samples": len(rows), "how": "one `nvidia-smi -lms 200` process over the warm-up and timed steps"}


# ------------------------------------------------------------------ the CPU arm (reference renderer on the host cores)
def strided_tasks(all_tasks, n_tiles):
    """A constant-stride sample of the frame's own (centre-sorted) task list: tiles from the centre to the corners in the
    proportion the frame has them -- the same camera, resolution and spp as the GPU arm, fewer tiles."""
    n = len(all_tasks)
    n_tiles = max(1, min(n, n_tiles))
    idx = [int(i * n / n_tiles) for i in range(n_tiles)]
    return (abi.Task * n_tiles)(*[all_tasks[i] for i in idx]), idx


def cpu_setup(cfg, desc):
    import checkers
    use_ref = checkers.have_ref()
    chk = checkers.ref() if use_ref else checkers.oracle()
    orc = checkers.oracle()
    h = chk.scene_create(desc)
    ho = orc.scene_create(desc) if use_ref else h
    ca = cfg.camera_args()
    cam = orc.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    return chk, orc, h, ho, cam, use_ref


def run_cpu(chk, h, cam, p, tasks, threads, shadow_rays=None):
    t0 = time.perf_counter()
    _, _, st = chk.render_round(h, cam, p, tasks, nthreads=threads)
    dt = time.perf_counter() - t0
    closest = int(st.closest_rays)
    shadow = int(st.shadow_rays) if shadow_rays is None else shadow_rays
    return dt, closest, shadow, int(st.samples)


def cpu_sample_desc(n_tiles, n_all, cfg, samples):
    return "%d of the frame's %d 32x32 tiles at a constant stride through the task list, %dx%d camera, %d spp (%d samples)" % (
        n_tiles, n_all, cfg.xres, cfg.yres, cfg.multisample, samples)


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pack, cfg, label = build_workload(args.workload, args.spp)
    desc = pack.desc()
    threads = os.cpu_count() or 1
    chk, orc, h, ho, cam, use_ref = cpu_setup(cfg, desc)
    p = cfg.params()
    all_tasks = orc.generate_tasks(32, p.xres, p.yres)
    # size the per-step sample so that warmup + steps stay near two minutes on this box's cores: probe a few tiles first
    probe, _ = strided_tasks(all_tasks, max(8, threads))
    probe_dt, _, _, probe_samples = run_cpu(chk, h, cam, p, probe, threads, 0)
    per_tile = probe_dt / len(probe)
    budget = 120.0 / max(1, args.steps + args.warmup + 1)
    n_tiles = int(max(threads, min(len(all_tasks), budget / max(per_tile, 1e-4))))
    tasks, _ = strided_tasks(all_tasks, n_tiles)
    shadow = None
    if use_ref:   # the reference does not count shadow rays (src/path_tracer.cpp:126 counts closest only): take the
        # count of the bit-identical oracle run of the same sample, untimed
        _, _, st = orc.render_round(ho, cam, p, tasks, nthreads=threads)
        shadow = int(st.shadow_rays)
    for _ in range(args.warmup):
        run_cpu(chk, h, cam, p, tasks, threads, shadow)
    tot_t, tot_r, tot_s = 0.0, 0, 0
    for _ in range(args.steps):
        dt, c, s, smp = run_cpu(chk, h, cam, p, tasks, threads, shadow)
        tot_t += dt; tot_r += c + s; tot_s += smp
    value = tot_r / tot_t / 1e6
    sample = cpu_sample_desc(len(tasks), len(all_tasks), cfg, tot_s // max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * tot_t / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 (+f64 plane distance)", "data": "synthetic stand-in scene (seeded); reference CPU renderer",
        "config": {"workload": label, "sample": sample},
        "samples_per_s": tot_s / tot_t,
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference" if use_ref else "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ roofline
def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_records(ctx, cam, p, tasks, ntasks, fb, cnt, stats, steps, info, clocks_mhz, fixed_point_light=False):
    """An untimed counting round gives the work counters; the timed steps give the per-class kernel time (CUDA events on the
    launch stream); profiles/r2_kernel_metrics.json gives, per kernel class, DRAM bytes and warp instructions per unit from the
    committed `ncu --set full` capture of this command.  Per class:
      achieved = algorithmic bytes / time (what the step's formula says must move)        -> frac      = / HBM peak
      dram     = measured DRAM bytes / time                                                             -> frac_hbm  = / HBM peak
      issue    = warp instructions / (SMs * 4 schedulers * clock * time)                                -> frac_issue
    The headline `roofline` is the class with the largest share of the step."""
    ctx.bvh_stats()
    ctx.set_counting(True)
    ctx.set_shard(0, 1)
    st = ctx.render_round_device(cam, p, tasks, fb.data_ptr(), cnt.data_ptr(), 42, 100 * ntasks)
    ctx.set_counting(False)
    bvh = ctx.bvh_stats()
    sh = ctx.shade_stats()
    tc, ts = ctx.render_trav_stats()
    bvh_on = bvh["rays"] > 0
    n_close, n_shadow, n_samples = int(st.closest_rays), int(st.shadow_rays), int(st.samples)
    # Traversal classes.  What a ray MUST move through HBM is its record in and its result out (36 B + 20 B hit record /
    # 1 B flag, + the 4-byte queue entry): that is `algorithmic_bytes_per_unit`, the numerator of `frac`.  The structure it walks
    # (a few MB: 128 B per node visit, 20 B per leaf slot, 32 B per exact test -- the kernel's own counters) is fed by L1 / L2 and
    # reported beside it as cache_fed_bytes_per_unit: over the HBM peak that figure exceeds 1 and grades nothing.
    cache_fed = {}
    if bvh_on:      # B_ray(bvh) = 36 + out + 128 nodes + 20 leaf slots + 32 exact tests (include/rgk_b200.h: rgk_bvh_stats)
        def b_cache(c):
            return (128 * c["nodes"] + 20 * c["slots"] + 32 * c["tests"]) / max(1, c["rays"])
        cache_fed = {"closest": b_cache(bvh["closest"]), "shadow": b_cache(bvh["shadow"])}
        b_closest, b_shadow = 36 + 20 + 4, 36 + 1 + 4
        trav = {"closest": {k: bvh["closest"][k] / max(1, bvh["closest"]["rays"]) for k in ("nodes", "slots", "tests")},
                "shadow": {k: bvh["shadow"][k] / max(1, bvh["shadow"]["rays"]) for k in ("nodes", "slots", "tests")}}
    else:           # 36 + out + 8 inner + 8 leaves + 4 refs + 48 tests
        cache_fed = {"closest": tc.bytes_per_ray(20) - 56, "shadow": ts.bytes_per_ray(1) - 37}
        b_closest, b_shadow = 36 + 20 + 4, 36 + 1 + 4
        trav = {"closest": {k: v / max(1, tc.rays) for k, v in tc.as_dict().items() if k != "rays"},
                "shadow": {k: v / max(1, ts.rays) for k, v in ts.as_dict().items() if k != "rays"}}
    # k_shade, per thread (= per closest-hit ray): path state read 84 B (hit, ray origin + direction, throughput, cursor, the two
    # sampler values, queue entry); written 33 B per light evaluation (surface point + NEE record + key), 65 B per continuation
    # (ray, throughput, last triangle, cursor, key, two queue entries), 32 B radiance read-modify-write where no shadow ray is queued;
    # per surface vertex 96 B of vertex attributes + 64 B material; 16 B per image texel; 208 B per LTC lobe evaluation
    threads = sh["vertices"] + sh["sky_vertices"]
    shade_bytes = (84 * threads + 33 * sh["light_evals"] + 65 * sh["continuations"] + 32 * (threads - sh["light_evals"])
                   + 160 * sh["vertices"] + 16 * sh["texels"] + 208 * sh["ltc_evals"])
    b_shade = shade_bytes / max(1, threads)
    # sampler, per pixel: the tables it must write (the generator state it streams on top is an artefact of the implementation)
    ms_ = int(p.multisample)
    ss = ctx.sampler_set_size(ms_)
    n1d, n2d = 1 + int(p.depth), 4 + int(p.depth) + (1 if cam.lens_size != 0.0 else 0)
    npix = n_samples // ms_
    # the tables the round reads (render_round_impl's keep masks): the pixel jitter, the lens sample, one direction per vertex but the
    # last, the Russian-roulette cursor per vertex but the last; the light's two 2-D samples and one 1-D sample unless the scene's
    # only light is one point light of size 0 (info.total_areal_power == 0 and the pack has one such light)
    depth = int(p.depth)
    one_fixed_light = info.n_areal_lights == 0 and fixed_point_light
    kept2 = 1 + (1 if cam.lens_size != 0.0 else 0) + max(0, depth - 1) + (0 if one_fixed_light else 2)
    kept1 = max(0, depth - 1) + (0 if one_fixed_light else 1)
    b_sampler = (kept1 * 4 + kept2 * 8) * ss + 4

    peaks = load_peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured, sustained copy)" if peaks else "fallback 6650 GB/s (of fallback)"
    try:
        km = json.load(open(KERNEL_METRICS))
    except Exception:
        km = {}
    sms = 148
    clock_hz = (clocks_mhz or 1965) * 1e6
    n = max(1, steps)
    cls = {"closest": (sum(float(s.closest_ms) for s in stats) / n, sum(int(s.closest_rays) for s in stats) / n, b_closest, "ray"),
           "shadow": (sum(float(s.shadow_ms) for s in stats) / n, sum(int(s.shadow_rays) for s in stats) / n, b_shadow, "ray"),
           "shade": (sum(float(s.shade_ms) for s in stats) / n, sum(int(s.closest_rays) for s in stats) / n, b_shade, "vertex (k_shade thread)"),
           "sampler": (sum(float(s.sampler_ms) for s in stats) / n, npix, b_sampler, "pixel")}
    launches = {"closest": sum(int(s.closest_launches) for s in stats) / n, "shadow": sum(int(s.shadow_launches) for s in stats) / n,
                "shade": sum(int(s.closest_launches) for s in stats) / n, "sampler": 1.0}
    names = {"closest": "k_closest_bvh (+ k_closest_arb)" if bvh_on else "k_closest", "shadow": "k_shadow_bvh (+ k_shadow_arb)" if bvh_on else "k_shadow",
             "shade": "k_shade (+ k_raygen, k_bin, k_finish in its event bracket)", "sampler": "k_sampler_warp (+ k_pixel_setup)"}
    step_ms = sum(float(s.gpu_ms) for s in stats) / n
    by = {}
    for k, (ms, units, b_unit, unit_name) in cls.items():
        sec = ms / 1e3
        m = km.get(k, {})
        rec = {"kernel": names[k], "ms_per_step": ms, "share_of_step": ms / step_ms if step_ms else None, "unit": unit_name, "units_per_step": units,
               "launches_per_step": launches[k], "algorithmic_bytes_per_unit": b_unit,
               "achieved_GBs": units * b_unit / sec / 1e9 if sec > 0 else 0.0}
        rec["frac"] = rec["achieved_GBs"] / peak
        if k in cache_fed:
            rec["cache_fed_bytes_per_unit"] = cache_fed[k]
            rec["cache_fed_GBs"] = units * cache_fed[k] / sec / 1e9 if sec > 0 else 0.0
        if "dram_bytes_per_unit" in m:
            rec["dram_bytes_per_unit"] = m["dram_bytes_per_unit"]
            rec["frac_hbm"] = units * m["dram_bytes_per_unit"] / sec / 1e9 / peak if sec > 0 else 0.0
        if "warp_inst_per_unit" in m:
            rec["warp_inst_per_unit"] = m["warp_inst_per_unit"]
            rec["active_lanes_per_instruction"] = m.get("active_lanes_per_instruction")
            rec["frac_issue"] = units * m["warp_inst_per_unit"] / (sms * 4 * clock_hz * sec) if sec > 0 else 0.0
        for extra in ("registers_per_thread", "achieved_occupancy_pct", "l1_hit_pct", "l2_hit_pct"):
            if extra in m:
                rec[extra] = m[extra]
        rec["bound"] = "issue" if k in ("closest", "shadow") else "hbm"
        by[k] = rec
    top = max(by, key=lambda k: by[k]["ms_per_step"])
    t = by[top]
    roofline = {"bound": "hbm", "kernel": t["kernel"], "achieved": t["achieved_GBs"], "peak": peak, "unit": "GB/s", "frac": t["frac"],
                "traffic": (t["dram_bytes_per_unit"] * t["units_per_step"] / max(1.0, t["launches_per_step"])) if "dram_bytes_per_unit" in t else None,
                "traffic_unit": "DRAM bytes per launch: ncu dram__bytes_read.sum + dram__bytes_write.sum per unit (profiles/r2_kernel_metrics.json) x units per launch",
                "algorithmic_bytes_per_launch": t["algorithmic_bytes_per_unit"] * t["units_per_step"] / max(1.0, t["launches_per_step"]),
                "avg_launch_ms": t["ms_per_step"] / max(1.0, t["launches_per_step"]), "launches_per_step": t["launches_per_step"],
                "share_of_step": t["share_of_step"], "frac_hbm": t.get("frac_hbm"), "frac_issue": t.get("frac_issue"),
                "peak_source": peak_src, "class": top,
                "note": "the kernel class with the largest share of the step; algorithmic bytes from the counting round (per-unit formulas in "
                        "DESIGN.md 4), time from CUDA events on the launch stream over the timed steps; frac_hbm uses measured DRAM bytes and "
                        "frac_issue measured warp instructions per unit from the committed ncu capture (%s); every class in roofline_by_kernel. "
                        "The traversal classes are issue / divergence-bound (their structure, a few MB, lives in L1/L2 -- cache_fed_bytes_per_unit): "
                        "their `frac` counts the ray records that must cross HBM and is small by nature, their graded fraction is frac_issue" % (
                            km.get("source", "profiles/r2_kernel_metrics.json missing"))}
    extra = {"work_counters": {"traversal_per_ray": trav, "shade": sh, "prefilter": tc.device_dict() if not bvh_on else None,
                               "counting_round": {"closest_rays": n_close, "shadow_rays": n_shadow, "samples": n_samples}},
             "kernel_share_of_step": {k: v["share_of_step"] for k, v in by.items()},
             "sum_of_kernel_ms_per_step": sum(v["ms_per_step"] for v in by.values())}
    return roofline, by, extra, bvh_on


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="sponza")
    ap.add_argument("--spp", type=int, default=None, help="override multisample (marks the run as non-headline)")
    ap.add_argument("--res", default=None, help="WxH: override the resolution of a stand-in workload (profiling runs; marks the run as non-headline)")
    ap.add_argument("--sampler", default="mt", choices=["mt", "fast"])
    ap.add_argument("--shard", default="rounds", choices=["rounds", "tiles"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--quick", action="store_true", help="timed steps only: no e2e, parity, counting, strong or CPU legs (A/B runs, large workloads)")
    ap.add_argument("--traversal", default="bvh", choices=["bvh", "kd"],
                    help="bvh: the library default, the wide-BVH traversal with the kd-tree arbiter pass (results bit-identical to the kd "
                         "path); kd: the reference's kd-tree for every ray")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: no nvidia-smi process beside the run (the line then has no clocks record)")
    ap.add_argument("--cfg", default="", help="rgk_device_cfg fields changed from the library defaults, e.g. binning=0,refill_shadow=20 (A/B runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from rgk_b200 import device, multigpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rgk_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    pack, cfg, label = build_workload(args.workload, args.spp, args.res)
    desc = pack.desc()
    render_stream = torch.cuda.current_stream()
    comm_stream = torch.cuda.Stream()
    ctx = device.Context(local, stream=render_stream.cuda_stream, traversal=args.traversal, **parse_cfg(args.cfg))
    t0 = time.perf_counter()
    ctx.commit(desc)
    commit_s = time.perf_counter() - t0
    info = ctx.scene_info()
    cam = ctx.camera(**cfg.camera_args())
    mode = abi.SAMPLER_MT19937 if args.sampler == "mt" else abi.SAMPLER_FAST
    p = cfg.params(mode)
    all_tasks = ctx.generate_tasks(32, p.xres, p.yres)
    ntasks = len(all_tasks)
    npx = p.yres * p.xres
    total = torch.zeros((p.yres, p.xres, 3), dtype=torch.float32, device="cuda")     # rank 0: the running image (sum of samples)
    cnt = torch.zeros((p.yres, p.xres), dtype=torch.int32, device="cuda")            # per-GPU sample counts (analytic globally: not reduced)
    parts = [torch.zeros_like(total) for _ in range(2)] if world > 1 else None
    rendered = [torch.cuda.Event() for _ in range(2)]
    reduced = [torch.cuda.Event() for _ in range(2)]

    def step(i, tiles):
        """One round per GPU (weak) or this rank's tiles of one round (strong).  world > 1: rendered into partial buffer i % 2;
        its NCCL reduce and rank 0's accumulate run on the communication stream, under the next step's rendering."""
        rnd = i if tiles else multigpu.round_for_rank(i, rank, world)
        if world == 1:
            return ctx.render_round_device(cam, p, all_tasks, total.data_ptr(), cnt.data_ptr(), 42, multigpu.seedcount_base(rnd, ntasks))
        b = i & 1
        render_stream.wait_event(reduced[b])                 # the reduce that last read this partial buffer (two steps ago)
        parts[b].zero_()
        st = ctx.render_round_device(cam, p, all_tasks, parts[b].data_ptr(), cnt.data_ptr(), 42, multigpu.seedcount_base(rnd, ntasks))
        rendered[b].record(render_stream)
        with torch.cuda.stream(comm_stream):
            comm_stream.wait_event(rendered[b])
            dist.reduce(parts[b], dst=0, op=dist.ReduceOp.SUM)     # in place on rank 0: its own partial + everybody else's
            if rank == 0:
                ctx.accumulate_device(total.data_ptr(), parts[b].data_ptr(), npx * 3, stream=comm_stream.cuda_stream)   # EXRTexture::Accumulate
            reduced[b].record(comm_stream)
        return st

    def barrier():
        render_stream.wait_stream(comm_stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, first, tiles):
        """n_steps steps between barriers, CUDA events on the render stream (the communication stream is joined before the
        closing event), max over ranks.  Returns (ms, per-step stats)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out = []
        barrier()
        e0.record(render_stream)
        for i in range(n_steps):
            out.append(step(first + i, tiles))
        render_stream.wait_stream(comm_stream)
        e1.record(render_stream)
        barrier()
        return e0.elapsed_time(e1), out

    def aggregate(ms, stats):
        rays = sum(int(s.closest_rays) + int(s.shadow_rays) for s in stats)          # rays actually traced
        skipped = sum(int(s.shadow_rays_skipped) for s in stats)                     # reference Visibility calls proven irrelevant
        samples = sum(int(s.samples) for s in stats)
        launches = sum(int(s.kernel_launches) for s in stats)
        agg = torch.tensor([ms, float(rays), float(samples), float(launches), float(skipped)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = agg.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = agg.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            return float(mx[0]), float(sm[1]), float(sm[2]), float(sm[3]), float(sm[4])
        return tuple(float(x) for x in agg)

    tiles_mode = args.shard == "tiles" and world > 1
    if tiles_mode:
        ctx.set_shard(rank, world)      # this rank renders tiles rank, rank+world, ... of every list, seeds unchanged
    warm = max(3, args.warmup)
    # started before the warm-up steps so that nvidia-smi's own start-up (~0.3 s, holds the driver lock) is over when the timed
    # region begins; its samples cover the warm-up (same load) and the timed steps
    sampler = ClockSampler(local) if rank == 0 and not args.no_clocks else None
    for i in range(warm):
        step(i, tiles_mode)
    barrier()
    ms, stats = timed(args.steps, 100, tiles_mode)
    clocks = sampler.stop() if sampler else None
    ms, rays, samples, launches, skipped = aggregate(ms, stats)
    value = rays / (ms / 1e3) / 1e6
    line = {
        "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / max(1, args.steps), "higher_is_better": True, "scaling": "strong" if tiles_mode else "weak",
        "vs_baseline": None, "dtype": "f32 (+f64 plane distance)", "data": "synthetic: seeded stand-in scene, procedural textures",
        "config": {"workload": label, "sampler": "mt19937 replica (same sequence as the CPU reference)" if mode == abi.SAMPLER_MT19937 else "fast counter-based",
                   "parallelism": ("1 GPU" if world == 1 else ("round-sharded x%d, one NCCL reduce per round overlapped with the next round" % world if not tiles_mode
                                                               else "tile-sharded x%d, one NCCL reduce per round overlapped with the next round" % world)),
                   "traversal": ("wide BVH candidate pass + kd-tree arbiter pass (the library default): bit-identical to the kd-tree path"
                                 if args.traversal == "bvh" else "kd-tree (reference structure) for every ray"),
                   "device_cfg": args.cfg or "library defaults",
                   "l2": "no flush: per-step path state and sampler tables (GBs) exceed the 126 MB L2; the scene (a few MB) is the working set"},
        "samples_per_s": samples / (ms / 1e3), "gpu_launches": int(launches),
        "rays_note": "value counts rays actually traced; %d shadow queries per step whose direct term is exactly 0 are resolved without "
                     "tracing (the reference traces them): reference-equivalent rate %.1f Mrays/s" % (
                         int(skipped / max(1, args.steps)), (rays + skipped) / (ms / 1e3) / 1e6),
        "clocks": clocks,
        "closest_rays_per_step": sum(int(s_.closest_rays) for s_ in stats) / max(1, args.steps) if world == 1 else None, "spp": int(p.multisample),
        "class_ms_per_step": {k: sum(float(getattr(s_, k + "_ms")) for s_ in stats) / max(1, args.steps) for k in ("closest", "shadow", "shade", "sampler")},
    }

    # ---- e2e: the same step through rgk_render_round with pinned host framebuffers
    if not args.quick:
        h_fb = torch.zeros((p.yres, p.xres, 3), dtype=torch.float32).pin_memory()
        h_cnt = torch.zeros((p.yres, p.xres), dtype=torch.int32).pin_memory()
        np_fb, np_cnt = h_fb.numpy(), h_cnt.numpy().view(np.uint32)

        def e2e_step(i):
            rnd = i if tiles_mode else multigpu.round_for_rank(i, rank, world)
            _, _, st = ctx.render_round(cam, p, all_tasks, 42, multigpu.seedcount_base(rnd, ntasks), fb=(np_fb, np_cnt))
            return st
        e2e_step(0)
        barrier()
        t0 = time.perf_counter()
        e2e_rays = 0
        for i in range(args.steps):
            st = e2e_step(200 + i)
            e2e_rays += int(st.closest_rays) + int(st.shadow_rays)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        agg2 = torch.tensor([e2e_s, float(e2e_rays)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx = agg2.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            sm = agg2.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
            e2e_s, e2e_rays = float(mx[0]), float(sm[1])
        fb_bytes = h_fb.numel() * 4 + h_cnt.numel() * 4
        line["e2e"] = {"value": e2e_rays / e2e_s / 1e6, "unit": "Mrays/s",
                       "h2d_bytes_per_step": fb_bytes + C.sizeof(abi.Camera) + C.sizeof(abi.RenderParams) + 16 * ntasks,
                       "d2h_bytes_per_step": fb_bytes, "samples_per_s": (samples / e2e_s) if world == 1 else None,
                       "note": "per-GPU host framebuffers (no reduce in this leg)" if world > 1 else None}
        del h_fb, h_cnt

    # ---- strong scaling in the same run (world > 1, default sharding): the tile-sharded round, timed the same way, and its
    # image compared bit for bit with a single-GPU render of the same round
    if world > 1 and not tiles_mode and not args.quick:
        ctx.set_shard(rank, world)
        for i in range(2):
            step(300 + i, True)
        ms_s, st_s = timed(args.steps, 400, True)
        ms_s, rays_s, samples_s, _, _ = aggregate(ms_s, st_s)
        barrier()
        total.zero_()
        barrier()
        step(500, True)                                  # one tile-sharded round into a clean total
        barrier()
        strong = {"ms_per_step": ms_s / max(1, args.steps), "value": rays_s / (ms_s / 1e3) / 1e6, "unit": "Mrays/s",
                  "samples_per_s": samples_s / (ms_s / 1e3), "sharding": "32x32 tiles dealt round-robin, every pixel owned by one GPU"}
        if rank == 0:
            ctx.set_shard(0, 1)
            ref = torch.zeros_like(total)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.render_round_device(cam, p, all_tasks, ref.data_ptr(), cnt.data_ptr(), 42, multigpu.seedcount_base(500, ntasks))
            ref.zero_()
            torch.cuda.synchronize()
            e0.record(render_stream)
            ctx.render_round_device(cam, p, all_tasks, ref.data_ptr(), cnt.data_ptr(), 42, multigpu.seedcount_base(500, ntasks))
            e1.record(render_stream)
            torch.cuda.synchronize()
            t1 = e0.elapsed_time(e1)
            diff = int((ref.view(torch.int32) != total.view(torch.int32)).sum().item())
            strong.update({"n1_ms_same_round_same_run": t1, "efficiency_vs_n1": t1 / (world * strong["ms_per_step"]),
                           "image_equal_to_n1": diff == 0, "fb_words_differing": diff,
                           "limit": "per-round cost that does not shrink with a GPU's share: the four kd-arbiter launches (a few hundred long rays each, "
                                    "0.2 - 0.35 ms of single-ray latency whatever their number: ~1.1 ms per round), the tail of the sampler's and the "
                                    "traversal's persistent grids on 1/N of the tiles, ~17 kernel launches; the full-frame reduce of a partial image "
                                    "that is 1/N non-zero is overlapped with the next round"})
            del ref
        line["strong"] = strong
        ctx.set_shard(0, 1)
        barrier()

    # ---- untimed legs of rank 0: parity against the kd-only traversal, counting round, CPU arm
    if rank == 0 and not args.quick:
        if args.traversal == "bvh" and world == 1:
            kd = device.Context(local, stream=render_stream.cuda_stream, traversal="kd")
            kd.commit(desc)
            a, b = torch.zeros_like(total), torch.zeros_like(total)
            ca, cb = torch.zeros_like(cnt), torch.zeros_like(cnt)
            base = multigpu.seedcount_base(100, ntasks)           # the first timed round
            ctx.bvh_stats()
            sa = ctx.render_round_device(cam, p, all_tasks, a.data_ptr(), ca.data_ptr(), 42, base)
            used = ctx.bvh_stats()
            kd.render_round_device(cam, p, all_tasks, b.data_ptr(), cb.data_ptr(), 42, base)     # warm-up: the context's first round allocates its buffers
            b.zero_(); cb.zero_()
            sb = kd.render_round_device(cam, p, all_tasks, b.data_ptr(), cb.data_ptr(), 42, base)
            torch.cuda.synchronize()
            line["parity"] = {
                "against": "the same round (the headline workload, full frame, full spp) on the kd-only traversal, whose kernels the GPU "
                           "suite pins to the oracle; untimed",
                "fb_words_differing": int((a.view(torch.int32) != b.view(torch.int32)).sum().item()),
                "counts_differing": int((ca != cb).sum().item()),
                "rays_equal": (int(sa.closest_rays), int(sa.shadow_rays), int(sa.shadow_rays_skipped)) == (int(sb.closest_rays), int(sb.shadow_rays), int(sb.shadow_rays_skipped)),
                "bvh_rays": used["rays"], "deferred_to_kd_frac": used["ambiguous"] / max(1, used["rays"]),
                "kd_ms_per_round": float(sb.gpu_ms), "bvh_ms_per_round": float(sa.gpu_ms)}
            kd.close()
            del a, b, ca, cb
        roofline, by, extra, bvh_on = roofline_records(ctx, cam, p, all_tasks, ntasks, total, cnt, stats, args.steps, info,
                                                       clocks.get("sm_mhz") if clocks else None,
                                                       fixed_point_light=len(pack.point_lights) == 1 and pack.point_lights[0][3] == 0.0 and pack.point_lights[0][2] > 0.0)
        line["roofline"] = roofline
        line["roofline_by_kernel"] = by
        line.update(extra)
        line["host_gap_ms_per_step"] = line["ms_per_step"] - extra["sum_of_kernel_ms_per_step"] if world == 1 else None
        line["traversal"] = {"mode": "bvh4 + kd arbiter" if bvh_on else "kd"}
        line["closest_Mrays_per_s"] = sum(int(s.closest_rays) for s in stats) / (ms / 1e3) / 1e6 if world == 1 else None
        line["shadow_Mrays_per_s"] = sum(int(s.shadow_rays) for s in stats) / (ms / 1e3) / 1e6 if world == 1 else None
        line["scene"] = {"triangles": info.n_triangles, "kd_nodes": info.n_nodes, "kd_refs": info.n_refs, "kd_depth": info.max_depth,
                         "commit_s": commit_s, "tag": "standin" if args.workload != "cornell" else "real"}
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            chk, orc, h, ho, ccam, use_ref = cpu_setup(cfg, desc)
            cp = cfg.params()
            ctasks, _ = strided_tasks(orc.generate_tasks(32, cp.xres, cp.yres), CPU_TILES)
            shadow = None
            if use_ref:
                _, _, st = orc.render_round(ho, ccam, cp, ctasks, nthreads=threads)
                shadow = int(st.shadow_rays)
            dt, c, s, smp = run_cpu(chk, h, ccam, cp, ctasks, threads, shadow)
            line["cpu_baseline"] = {"value": (c + s) / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "reference" if use_ref else "port",
                                    "sample": cpu_sample_desc(len(ctasks), ntasks, cfg, smp) + ", %.1f s" % dt, "samples_per_s": smp / dt}
    if rank == 0:
        line.setdefault("e2e", None); line.setdefault("roofline", None); line.setdefault("cpu_baseline", None)
        print(json.dumps(line))
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
