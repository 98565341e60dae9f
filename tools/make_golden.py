"""Generate tests/golden/*.npz from the REFERENCE build (oracle/_ref/librgk_ref.so = the unmodified RGKrt
sources compiled in this container against the GLM-formula shim).  The reference ships no tests or golden
vectors, so these fixtures are outputs of the reference itself; they pin the CPU oracle
(tests/test_oracle_golden.py) on machines where /root/reference does not exist (the GPU box).
Run in the build container:  python tools/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import checkers  # noqa: E402
import raybatches  # noqa: E402
from rgk_b200 import scenes, abi  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
R = checkers.ref()


def cam_of(chk, cfg):
    ca = cfg.camera_args()
    return chk.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])


def info_dict(info):
    return dict(epsilon=np.float32(info.epsilon), bbox=np.array(list(info.bbox), np.float32), n_nodes=info.n_nodes, n_refs=info.n_refs,
                max_depth=info.max_depth, total_point_power=np.float32(info.total_point_power), total_areal_power=np.float32(info.total_areal_power),
                n_areal_lights=info.n_areal_lights)


# ---- 1. Cornell box: tree, planes, camera, rays, hits
pack, cfg = scenes.load_builtin("cornell-box", width=64, height=64, multisample=4)
desc = pack.desc()
h = R.scene_create(desc)
nodes, refs = R.scene_kdtree(h)
info = R.scene_info(h)
cam = cam_of(R, cfg)
rays = raybatches.primary(R, cam, 64, 64)
hits = R.trace_closest(h, rays)
planes = R.scene_planes(h)
brays, ign = raybatches.bounce(rays, hits, planes[:, :3], info.epsilon)
bhits = R.trace_closest(h, brays, ign)
bhits_noign = R.trace_closest(h, brays)
sa, sb = raybatches.shadow_segments(rays, hits, (-0.005, 1.97, -0.03))
vis = R.trace_shadow(h, sa, sb)
np.savez_compressed(os.path.join(OUT, "cornell_geometry.npz"), nodes=nodes, refs=refs, planes=planes, camera=np.frombuffer(bytes(cam), np.uint8),
                    rays=rays, hits=hits, brays=brays, ign=ign, bhits=bhits, bhits_noign=bhits_noign, sa=sa, sb=sb, vis=vis, **info_dict(info))
tasks = R.generate_tasks(32, 64, 64)
p = cfg.params()
p.depth = 40
fb, cnt, st = R.render_round(h, cam, p, tasks, nthreads=4)
tl = np.array([[t.x1, t.x2, t.y1, t.y2] for t in R.generate_tasks(32, 1920, 1080)], np.uint32)
np.savez_compressed(os.path.join(OUT, "cornell_render.npz"), fb=fb, cnt=cnt, closest_rays=int(st.closest_rays),
                    tasks_1080p=tl, tasks_64=np.array([[t.x1, t.x2, t.y1, t.y2] for t in tasks], np.uint32))
# bidirectional mode (reverse = 3: light path, camera splats, vertex connections), one worker thread so that the
# accumulation order of the per-task buffers is defined
p.depth, p.reverse = 6, 3
fbr, cntr, str_ = R.render_round(h, cam, p, tasks, nthreads=1)
np.savez_compressed(os.path.join(OUT, "cornell_reverse.npz"), fb=fbr, cnt=cntr, closest_rays=int(str_.closest_rays))

# ---- 2. StratifiedSampler tables
sam = {}
seeds = np.array([42 + 0x42424242, 7, 0, 0xFFFFFFFF], np.uint32)
for ms in (1, 4, 16, 40):
    t1, t2 = R.sampler_tables(seeds, ms, 64, 64)
    sam[f"t1_{ms}"], sam[f"t2_{ms}"] = t1, t2
    sam[f"set_size_{ms}"] = R.sampler_set_size(ms)
sam["set_sizes"] = np.array([[m, R.sampler_set_size(m)] for m in (1, 2, 16, 40, 50, 64, 256, 400, 512, 1024)], np.uint32)
np.savez_compressed(os.path.join(OUT, "sampler.npz"), seeds=seeds, **sam)

# ---- 3. material zoo: unit probes of every BxDF, textures, lights, sky + a small render
pack, cfg = scenes.material_zoo(width=48, height=32, multisample=4, lens=0.05)
desc = pack.desc()
h = R.scene_create(desc)
rng = np.random.default_rng(2024)
n = 96
zoo = {}
def unit(v):
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
Vi = unit(rng.normal(size=(n, 3))); Vi[: n // 2, 2] = np.abs(Vi[: n // 2, 2])
Vr = unit(rng.normal(size=(n, 3))); Vr[: n // 2, 2] = np.abs(Vr[: n // 2, 2])
# exact mirror / refraction partners so that the delta BxDFs return non-zero values too
Vr[:8] = Vi[:8] * np.array([-1, -1, 1], np.float32)
Vr[8:16] = -Vi[8:16]
uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
smp = rng.random((n, 2), dtype=np.float32)
zoo.update(Vi=Vi, Vr=Vr, uv=uv, smp=smp)
for mi in range(desc.n_materials):
    zoo[f"sample_{mi}"] = R.bxdf_sample(h, mi, Vi, uv, smp)
    zoo[f"value_{mi}"] = R.bxdf_value(h, mi, Vi, Vr, uv)
tuv = np.concatenate([rng.uniform(-2, 3, (200, 2)), np.array([[0, 0], [1, 1], [0.999999, 0.5], [1e-7, 1e-7], [-1e-7, 0.25], [0.5, -0.0], [0.0078125, 0.0078125],
                                                             [0.00390625, 0.99609375], [1.0 - 1.0 / 128, 0.5], [0.5, 1.0 / 128]])]).astype(np.float32)
zoo["tuv"] = tuv
for ti in range(desc.n_textures):
    zoo[f"tex_{ti}"] = R.texture_fetch(h, ti, tuv)
ls = rng.random((128, 5), dtype=np.float32)
zoo["light_samples"], zoo["lights"] = ls, R.random_light(h, ls)
dirs = unit(rng.normal(size=(128, 3)))
zoo["sky_dirs"], zoo["sky"] = dirs, R.sky(h, dirs)
cam = cam_of(R, cfg)
tasks = R.generate_tasks(32, 48, 32)
fb, cnt, st = R.render_round(h, cam, cfg.params(), tasks, nthreads=4)
zoo.update(fb=fb, cnt=cnt, closest_rays=int(st.closest_rays), camera=np.frombuffer(bytes(cam), np.uint8), **info_dict(R.scene_info(h)))
zn, zr = R.scene_kdtree(h)
zoo.update(nodes=zn, refs=zr)
np.savez_compressed(os.path.join(OUT, "zoo.npz"), **zoo)

# ---- 4. StratifiedSampler tables of the large set sizes (BASELINE C3 256 spp, C5 512 -> 529): the sets the device shuffles in
# place in global memory (k_sampler_mt<false>); a few dims of two seeds keep the file small
big = {}
seeds2 = np.array([42 + 0x42424242, 0x9E3779B9], np.uint32)
for ms in (256, 512):
    t1, t2 = R.sampler_tables(seeds2, ms, 3, 4)
    big[f"t1_{ms}"], big[f"t2_{ms}"] = t1, t2
np.savez_compressed(os.path.join(OUT, "sampler_large.npz"), seeds=seeds2, **big)

# ---- 5. a scene with a non-empty thinglass set (src/main.cpp:212): FindIntersectKdOtherThanWithThinglass / VisibilityWithThinglass
pack, cfg = scenes.load_builtin("cornell-box", width=64, height=64, multisample=4)
pack.thinglass = 1
desc = pack.desc()
h = R.scene_create(desc)
cam = cam_of(R, cfg)
rays = raybatches.primary(R, cam, 64, 64, jitter_seed=3)
hits = R.trace_closest(h, rays)
info = R.scene_info(h)
brays, ign = raybatches.bounce(rays, hits, R.scene_planes(h)[:, :3], info.epsilon, seed=8)
bhits = R.trace_closest(h, brays, ign)
sa, sb = raybatches.shadow_segments(brays, bhits, (-0.005, 1.97, -0.03))
vis = R.trace_shadow(h, sa, sb)
p = cfg.params(); p.depth = 40
tasks = R.generate_tasks(32, 64, 64)
fb, cnt, st = R.render_round(h, cam, p, tasks, nthreads=4)
np.savez_compressed(os.path.join(OUT, "cornell_thinglass.npz"), brays=brays, ign=ign, bhits=bhits, sa=sa, sb=sb, vis=vis, fb=fb, cnt=cnt,
                    closest_rays=int(st.closest_rays))
for f in sorted(os.listdir(OUT)):
    print(f, os.path.getsize(os.path.join(OUT, f)))
