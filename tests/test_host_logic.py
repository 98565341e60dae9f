"""Host-side logic that needs no GPU: the JSON scene config reader, the scene pack, sharding arithmetic, and the
product's host half of Scene::Commit (rgk_host_scene_*: planes, lights, kd-tree build incl. the forked build) against
the oracle's arrays."""
import json
import os

import numpy as np
import pytest

from rgk_b200 import abi, scene, scenes


def test_json_config_with_comments_and_defaults(tmp_path):
    cfgd = scenes.cornell_box(width=32, height=16, multisample=2)
    text = json.dumps(cfgd, indent=1).replace('"rounds": 1,', '"rounds": 1, // one round\n /* block\n comment */')
    path = tmp_path / "c.json"
    path.write_text(text)
    pack, cfg = scene.load_json_config(str(path))
    assert (cfg.xres, cfg.yres, cfg.multisample, cfg.recursion_level) == (32, 16, 2, 40)
    assert cfg.russian == pytest.approx(0.74) and cfg.clamp == 20.0 and cfg.bumpmap_scale == 1.0 and cfg.reverse == 0
    assert pack.n_triangles == 36 and len(pack.materials) == 4 and len(pack.meshes) == 9
    # defaults of ConfigJSON (src/config.cpp:295-301)
    d = dict(cfgd); [d.pop(k) for k in ("russian", "clamp", "recursion-max", "multisample")]
    _, cfg2 = scene.load_config(d)
    assert cfg2.russian == pytest.approx(0.74) and cfg2.clamp == 10000000.0 and cfg2.recursion_level == 40 and cfg2.multisample == 1


def test_config_errors_mirror_reference_exceptions():
    d = scenes.cornell_box()
    bad = dict(d); bad.pop("camera")
    with pytest.raises(scene.ConfigFileException):
        scene.load_config(bad)
    bad = dict(d); bad["render-time"] = 3
    with pytest.raises(scene.ConfigFileException):
        scene.load_config(bad)
    bad = json.loads(json.dumps(d)); bad["scene"][0]["material"] = "nope"
    with pytest.raises(ValueError, match="was not defined"):
        scene.load_config(bad)
    bad = json.loads(json.dumps(d)); bad["materials"][0]["brdf"] = "cooktorr"
    with pytest.raises(scene.ConfigFileException, match="Unsupported BRDF"):
        scene.load_config(bad)


def test_output_scale_key(tmp_path):
    """"output-scale": a number, or "auto" (src/config.cpp:303-312); it travels through the RGKPACK1 file to the C++ drivers."""
    d = scenes.cornell_box(width=32, height=16, multisample=1)
    assert scene.load_config(d)[1].output_scale == -1.0                     # absent: auto
    for given, want in ((0.5, 0.5), (2, 2.0), ("auto", -1.0)):
        dd = dict(d); dd["output-scale"] = given
        pack, cfg = scene.load_config(dd)
        assert cfg.output_scale == want
        path = str(tmp_path / "s.rgkpack")
        pack.save(path, cfg)
        assert scene.load_pack(path)[1].output_scale == want
    for bad_value in ("manual", [1.0], True, None):
        dd = dict(d); dd["output-scale"] = bad_value
        with pytest.raises(scene.ConfigFileException, match="output-scale"):
            scene.load_config(dd)


def test_color255_and_broadcast_vectors():
    d = scenes.cornell_box()
    d["materials"][0] = {"name": "LeftWall", "diffuse255": [255, 0, 127.5], "brdf": "diffuse"}
    d["sky"] = {"color": 0.25, "intensity": 2.0}
    d["lights"] = [{"position": [0, 1, 0], "intensity": 3.0, "color255": 255}]
    pack, _ = scene.load_config(d)
    tex = pack.textures[pack.materials[0]["tex_diffuse"]]
    assert tex[0] == "solid" and tex[1] == pytest.approx((1.0, 0.0, 0.5))
    assert pack.sky["color"] == (0.25, 0.25, 0.25) and pack.sky["intensity"] == 2.0
    assert pack.point_lights[0][1] == (1.0, 1.0, 1.0)


def test_scene_desc_layout(cornell):
    pack, cfg, desc = cornell
    assert desc.n_triangles == 36 and desc.n_vertices == 108 and desc.n_meshes == 9
    m = [desc.meshes[i] for i in range(desc.n_meshes)]
    assert [x.first_triangle for x in m] == [0, 2, 4, 6, 8, 10, 22, 34, 35] and sum(x.n_triangles for x in m) == 36
    pos = np.ctypeslib.as_array(desc.positions, (108 * 3,))
    assert not np.any(np.signbit(pos) & (pos == 0))          # -0.0 canonicalised like the reference's transform does
    assert desc.materials[3].emission[0] == 17.0
    ltc = scene.load_ltc_tables()
    assert ltc["ggx_M"].shape == (4096, 9) and ltc["beckmann_amp"].shape == (4096,)


def test_primitives_match_reference_tables():
    """planeY / trigY / cube as in src/primitives.cpp:168-228 (spot checks of the generated tables)."""
    P, N, UV, T = scene.primitive_data("plane")
    assert P.tolist() == [[1, 0, 1], [1, 0, -1], [-1, 0, 1], [-1, 0, -1], [-1, 0, 1], [1, 0, -1]]
    assert UV.tolist() == [[1, 1], [1, 0], [0, 1], [0, 0], [0, 1], [1, 0]] and N[0].tolist() == [0, 1, 0] and T[0].tolist() == [0, 0, 1]
    P, N, UV, T = scene.primitive_data("cube")
    assert len(P) == 36
    assert P[18].tolist() == [-1, -1, 1] and N[18].tolist() == [0, -1, 0] and UV[18].tolist() == [1, 1] and T[18].tolist() == [1, 0, 0]
    assert P[25].tolist() == [-1, 1, 1] and N[25].tolist() == [0, 0, 1] and UV[25].tolist() == [1, 0] and T[25].tolist() == [0, 1, 0]
    assert P[35].tolist() == [-1, 1, -1] and N[35].tolist() == [0, 0, -1]
    assert len(scene.primitive_data("tri")[0]) == 3


def test_camera_args_fov_and_focal():
    _, cfg = scenes.load_builtin("cornell-box", width=200, height=100)
    ca = cfg.camera_args()
    assert ca["xview"] == pytest.approx(2 * np.tan(np.float32(19.5 * 0.0174533) / 2), rel=1e-6)
    assert ca["yview"] == pytest.approx(ca["xview"] * 100 / 200, rel=1e-6)
    cfg.camera = {"position": [0, 0, 5], "lookat": [0, 0, 0], "focal": 1.6}
    ca = cfg.camera_args()
    assert ca["yview"] == pytest.approx(1.6) and ca["xview"] == pytest.approx(1.6 * 2)


def _host_tree(desc, threads):
    from rgk_b200 import device
    hs = device.HostScene(desc, build_threads=threads)
    info, (nodes, refs), (planes, rec) = hs.info(), hs.kdtree(), hs.records()
    hs.close()
    return info, nodes, refs, planes, rec


def test_host_scene_commit_matches_oracle_cornell(cornell, oracle):
    """rgk_host_scene_*: the product's host half of Scene::Commit (planes, epsilon, bbox, kd-tree) against the
    oracle's restatement of src/scene.cpp:294-657 -- byte-identical arrays."""
    pack, cfg, desc = cornell
    info, nodes, refs, planes, rec = _host_tree(desc, 1)
    ho = oracle.scene_create(desc)
    oi = oracle.scene_info(ho)
    on, orf = oracle.scene_kdtree(ho)
    assert nodes.tobytes() == on.tobytes() and refs.tobytes() == orf.tobytes()
    assert planes.tobytes() == oracle.scene_planes(ho).tobytes()
    assert info.epsilon == oi.epsilon and list(info.bbox) == list(oi.bbox) and info.max_depth == oi.max_depth
    assert info.n_areal_lights == oi.n_areal_lights == 2       # the lamp is two `tri` primitives = two ArealLights
    assert np.array_equal(rec[:, :4], planes)                  # record word 0 = the plane
    oracle.scene_destroy(ho)


def test_forked_kdtree_build_is_byte_identical_to_sequential():
    """the host-parallel kd-tree build (right subtrees built by other threads and stitched in) must
    give the arrays of the sequential reference procedure, whatever the thread count."""
    from rgk_b200 import standin
    pack, cfg = standin.sponza(width=64, height=64, multisample=1, target_tris=66000)
    desc = pack.desc()
    i1, n1, r1, p1, t1 = _host_tree(desc, 1)
    for threads in (2, 5, 16):
        i2, n2, r2, p2, t2 = _host_tree(desc, threads)
        assert n1.tobytes() == n2.tobytes() and r1.tobytes() == r2.tobytes(), threads
        assert i1.max_depth == i2.max_depth and p1.tobytes() == p2.tobytes() and t1.tobytes() == t2.tobytes()


def test_host_scene_rejects_bad_input():
    """Load-time errors of the reference (ConfigFileException / std::runtime_error) become RGK_ERR_INVALID + text."""
    from rgk_b200 import device
    pack, cfg = scenes.load_builtin("cornell-box")
    bad = pack.desc()
    bad.n_meshes = 8
    with pytest.raises(device.RgkError, match="mesh ranges"):
        device.HostScene(bad)
    bad = pack.desc()
    bad.materials[0].bxdf = 99
    with pytest.raises(device.RgkError, match="Unsupported BRDF"):
        device.HostScene(bad)


def test_json_reader_is_as_lenient_as_jsoncpp():
    """Comments, trailing commas and numbers with leading zeros (scenes/conference.json:16 has `000.0`)."""
    t = scene._strip_comments('{"a": [800.0, 400.0,  000.0, -007.5, 10.05, 100,], "s": "x 000.0 // y", /* c */ "b": 0.5, "d": 00, // e\n}')
    assert json.loads(t) == {"a": [800.0, 400.0, 0.0, -7.5, 10.05, 100], "s": "x 000.0 // y", "b": 0.5, "d": 0}
