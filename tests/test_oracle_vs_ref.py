"""Where the reference itself is built (oracle/_ref, only in the container that has /root/reference): the oracle
restatement against the reference on fresh, larger inputs than the committed fixtures.  Skipped elsewhere."""
import numpy as np
import pytest

import raybatches
from rgk_b200 import scenes, standin


def _cam(chk, cfg):
    ca = cfg.camera_args()
    return chk.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])


def test_cornell_full_resolution_hits(oracle, ref, cornell):
    pack, cfg, desc = cornell
    hr, ho = ref.scene_create(desc), oracle.scene_create(desc)
    assert all(np.array_equal(a, b) for a, b in zip(ref.scene_kdtree(hr), oracle.scene_kdtree(ho)))
    rays = raybatches.primary(ref, _cam(ref, cfg), 256, 256, jitter_seed=1)
    assert rays.tobytes() == raybatches.primary(oracle, _cam(oracle, cfg), 256, 256, jitter_seed=1).tobytes()
    assert ref.trace_closest(hr, rays).tobytes() == oracle.trace_closest(ho, rays).tobytes()


def test_standin_scene_tree_and_hits(oracle, ref):
    """A 70 k-triangle scene: the SAH build (std::sort per node, retries, depth cap) and incoherent rays."""
    pack, cfg = standin.sponza(width=160, height=90, multisample=1)
    desc = pack.desc()
    hr, ho = ref.scene_create(desc), oracle.scene_create(desc)
    assert all(np.array_equal(a, b) for a, b in zip(ref.scene_kdtree(hr), oracle.scene_kdtree(ho)))
    rays = raybatches.primary(ref, _cam(ref, cfg), 160, 90, jitter_seed=2)
    hits = ref.trace_closest(hr, rays)
    assert hits.tobytes() == oracle.trace_closest(ho, rays).tobytes()
    brays, ign = raybatches.bounce(rays, hits, ref.scene_planes(hr)[:, :3], ref.scene_info(hr).epsilon)
    assert ref.trace_closest(hr, brays, ign).tobytes() == oracle.trace_closest(ho, brays, ign).tobytes()
    a, b = raybatches.shadow_segments(rays, hits, (-16.0, 100.0, -10.0))
    assert np.array_equal(ref.trace_shadow(hr, a, b), oracle.trace_shadow(ho, a, b))


def test_render_rounds_bit_identical(oracle, ref):
    for pack, cfg in (scenes.load_builtin("cornell-box", width=48, height=40, multisample=9),
                      scenes.material_zoo(width=64, height=40, multisample=4, lens=0.03),
                      standin.sponza(width=64, height=36, multisample=4)):
        desc = pack.desc()
        hr, ho = ref.scene_create(desc), oracle.scene_create(desc)
        cam = _cam(oracle, cfg)
        p = cfg.params()
        tasks = oracle.generate_tasks(32, p.xres, p.yres)
        fr, cr, sr = ref.render_round(hr, cam, p, tasks, seedcount_base=5, nthreads=4)
        fo, co, so = oracle.render_round(ho, cam, p, tasks, seedcount_base=5, nthreads=4)
        assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32)) and np.array_equal(cr, co)
        assert int(sr.closest_rays) == int(so.closest_rays)


def test_bidirectional_rounds_bit_identical(oracle, ref):
    """reverse > 0 with one worker thread (with more, the reference's own accumulation order is a race)."""
    for (pack, cfg), rev in ((scenes.load_builtin("cornell-box", width=48, height=32, multisample=4), 2),
                             (scenes.load_builtin("cornell-box", width=48, height=32, multisample=4), 4),
                             (scenes.material_zoo(width=48, height=32, multisample=4), 3)):
        cfg.reverse = rev
        desc = pack.desc()
        hr, ho = ref.scene_create(desc), oracle.scene_create(desc)
        cam = _cam(oracle, cfg)
        p = cfg.params()
        tasks = oracle.generate_tasks(32, p.xres, p.yres)
        fr, cr, sr = ref.render_round(hr, cam, p, tasks, nthreads=1)
        fo, co, so = oracle.render_round(ho, cam, p, tasks, nthreads=1)
        assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32)) and np.array_equal(cr, co)
        assert int(sr.closest_rays) == int(so.closest_rays)


def test_every_loadable_reference_scene_file(oracle, ref):
    """Every scene file of the reference whose assets ship with it (20 of them: primitives, OBJ meshes with imported LTC
    materials, textures, bump maps, reverse 2-4, recursion up to 40) goes through the JSON reader and the asset loaders and
    renders bit-identically on the oracle and on the reference build; the rest fail for the reason upstream would give
    (missing model / envmap files, a material without "brdf", an LTC material without roughness)."""
    import glob
    import os
    import warnings
    from rgk_b200 import scene
    rendered, reasons = [], {}
    for path in sorted(glob.glob("/root/reference/scenes/*.json")):
        name = os.path.basename(path)
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                pack, cfg = scene.load_json_config(path, overrides={"output-width": 32, "output-height": 24, "multisample": 2})
        except (scene.ConfigFileException, FileNotFoundError) as e:
            reasons[name] = str(e)
            continue
        desc = pack.desc()
        hr, ho = ref.scene_create(desc), oracle.scene_create(desc)
        cam = _cam(oracle, cfg)
        p = cfg.params()
        tasks = oracle.generate_tasks(32, p.xres, p.yres)
        fr, cr, sr = ref.render_round(hr, cam, p, tasks, nthreads=1)
        fo, co, so = oracle.render_round(ho, cam, p, tasks, nthreads=1)
        assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32)) and np.array_equal(cr, co), name
        assert int(sr.closest_rays) == int(so.closest_rays), name
        rendered.append((name, cfg.reverse))
    if not rendered and not reasons:
        pytest.skip("no reference scenes here")
    assert len(rendered) >= 20 and sum(1 for _, r in rendered if r > 0) >= 5
    for name, why in reasons.items():
        assert ("Unable to find model file" in why or "file does not exist" in why or 'Required value "brdf"' in why
                or '"roughness" or "exponent"' in why), (name, why)


def test_integration_bridge_describes_the_reference_scene(ref, oracle):
    """integration/gpu_bridge.hpp (the binding a RGKrt maintainer adds) compiled against the reference's own headers:
    RgkGpuBridge::Describe turns the reference's Scene back into the scene pack it was loaded from, and the library's
    host commit of that description yields the reference's own flattened kd-tree."""
    import ctypes as C
    from rgk_b200 import abi, device
    lib = ref.lib
    lib.rgkref_bridge_describe.restype = C.c_void_p
    lib.rgkref_bridge_describe.argtypes = [C.c_void_p]
    lib.rgkref_bridge_desc.restype = C.POINTER(abi.SceneDesc)
    lib.rgkref_bridge_desc.argtypes = [C.c_void_p]
    lib.rgkref_bridge_destroy.argtypes = [C.c_void_p]

    def arr(ptr, n, dt=np.float32):
        return np.ctypeslib.as_array(ptr, (n,)).copy() if n else np.zeros(0, dt)

    def tex_key(d, i):
        if i < 0:
            return None
        t = d.textures[i]
        if t.kind == 0:
            return ("solid", tuple(t.color))
        return ("image", t.width, t.height, arr(t.texels, 3 * t.width * t.height).tobytes())

    for pack, cfg in (scenes.load_builtin("cornell-box", width=32, height=32, multisample=1),
                      scenes.material_zoo(width=32, height=32, multisample=1),
                      standin.sponza(width=32, height=32, multisample=1, target_tris=6000)):
        d1 = pack.desc()
        hr = ref.scene_create(d1)
        b = lib.rgkref_bridge_describe(hr)
        assert b
        d2 = lib.rgkref_bridge_desc(b).contents
        assert (d2.n_vertices, d2.n_triangles, d2.n_materials, d2.n_point_lights) == (d1.n_vertices, d1.n_triangles, d1.n_materials, d1.n_point_lights)
        for name, k in (("positions", 3), ("normals", 3), ("tangents", 3), ("texcoords", 2)):
            assert arr(getattr(d1, name), k * d1.n_vertices).tobytes() == arr(getattr(d2, name), k * d2.n_vertices).tobytes(), name
        assert np.array_equal(arr(d1.indices, 3 * d1.n_triangles, np.uint32), arr(d2.indices, 3 * d2.n_triangles, np.uint32))
        for i in range(d1.n_materials):
            a, c = d1.materials[i], d2.materials[i]
            assert (a.bxdf, a.no_russian, tuple(a.emission), a.mix_a, a.mix_b) == (c.bxdf, c.no_russian, tuple(c.emission), c.mix_a, c.mix_b), i
            used = {abi.BXDF_DIFFUSE: ("tex_diffuse",), abi.BXDF_MIX: (), abi.BXDF_TRANSPARENT: (), abi.BXDF_MIRROR: ("tex_color",), abi.BXDF_DIELECTRIC: ("tex_color",)}.get(
                a.bxdf, ("tex_color", "tex_diffuse") if a.bxdf in (abi.BXDF_LTC_GGX_DIFFUSE, abi.BXDF_LTC_BECKMANN_DIFFUSE) else ("tex_color",))
            for slot in used + ("tex_bump",):
                assert tex_key(d1, getattr(a, slot)) == tex_key(d2, getattr(c, slot)), (i, slot)
            if a.bxdf == abi.BXDF_MIX:
                assert a.amount == c.amount
            if a.bxdf == abi.BXDF_DIELECTRIC:
                assert a.ior == c.ior
            if a.bxdf >= abi.BXDF_LTC_BECKMANN:
                assert a.roughness == c.roughness
        for i in range(d1.n_point_lights):
            assert bytes(d1.point_lights[i]) == bytes(d2.point_lights[i])
        assert (d1.sky.mode, d1.sky.intensity, d1.sky.rotate) == (d2.sky.mode, d2.sky.intensity, d2.sky.rotate)
        assert tuple(d1.sky.color) == tuple(d2.sky.color) if d1.sky.mode == 0 else tex_key(d1, d1.sky.envmap) == tex_key(d2, d2.sky.envmap)
        for t in ("ltc_ggx", "ltc_beckmann"):
            assert arr(getattr(d1, t).M, 4096 * 9).tobytes() == arr(getattr(d2, t).M, 4096 * 9).tobytes()
            assert arr(getattr(d1, t).amplitude, 4096).tobytes() == arr(getattr(d2, t).amplitude, 4096).tobytes()
        # emissive meshes = areal lights: same grouping
        lights1 = [(m.first_triangle, m.n_triangles) for m in (d1.meshes[i] for i in range(d1.n_meshes)) if any(x > 0 for x in d1.materials[m.material].emission)]
        lights2 = [(m.first_triangle, m.n_triangles) for m in (d2.meshes[i] for i in range(d2.n_meshes)) if any(x > 0 for x in d2.materials[m.material].emission)]
        assert lights1 == lights2
        # the library's host commit of the bridged description == the reference's own flattened tree
        hs = device.HostScene(d2)
        n2, r2 = hs.kdtree()
        nr, rr = ref.scene_kdtree(hr)
        assert np.array_equal(n2, nr) and np.array_equal(r2, rr) and hs.info().n_areal_lights == ref.scene_info(hr).n_areal_lights
        hs.close()
        lib.rgkref_bridge_destroy(b)
