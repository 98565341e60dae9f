"""Where the reference itself is built (oracle/_ref, only in the container that has /root/reference): the oracle
restatement against the reference on fresh, larger inputs than the committed fixtures.  Skipped elsewhere."""
import numpy as np

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


def test_reference_scene_files_bidirectional(oracle, ref):
    """The reference's own bidirectional scenes (OBJ meshes, imported LTC materials, reverse 3-4) through the asset
    loaders: the oracle and the reference build render the same pack bit for bit."""
    import os
    import warnings
    from rgk_b200 import assets, scene
    for name in ("box2.json", "cb1.json", "box6.json"):
        path = os.path.join("/root/reference/scenes", name)
        if not os.path.exists(path):
            continue
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pack, cfg = scene.load_json_config(path, overrides={"output-width": 48, "output-height": 32, "multisample": 2},
                                               mesh_loader=assets.load_obj_into, texture_loader=assets.load_image)
        assert cfg.reverse >= 3
        desc = pack.desc()
        hr, ho = ref.scene_create(desc), oracle.scene_create(desc)
        cam = _cam(oracle, cfg)
        p = cfg.params()
        tasks = oracle.generate_tasks(32, p.xres, p.yres)
        fr, cr, _ = ref.render_round(hr, cam, p, tasks, nthreads=1)
        fo, co, _ = oracle.render_round(ho, cam, p, tasks, nthreads=1)
        assert np.array_equal(fr.view(np.uint32), fo.view(np.uint32)) and np.array_equal(cr, co)
