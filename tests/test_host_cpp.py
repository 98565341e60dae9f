"""The C++ host layer (include/rgk_b200_host.hpp: EXRTexture, RenderTask, Camera, PackFile, PathTracer, RenderDriver)
and the on-disk scene pack.  Host-only parts run here against the oracle / numpy; the rendering parts are GPU tests
that drive the example binary build/host/rgk_render and compare with the Python binding bit for bit."""
import json
import os
import subprocess

import numpy as np
import pytest

from rgk_b200 import scene, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "build", "host")


SOURCES = {"rgk_render": os.path.join(ROOT, "rgk_b200", "host", "rgk_render.cpp"), "host_checks": os.path.join(ROOT, "tests", "host_cpp", "host_checks.cpp"),
           "rgk_render_multi": os.path.join(ROOT, "rgk_b200", "host", "rgk_render_multi.cpp")}


def _need(exe):
    """The host programs are built by __graft_entry__.build(); if the build directory did not travel with the snapshot
    they are compiled here against the in-tree library (g++ only, no nvcc needed)."""
    path = os.path.join(HOST, exe)
    if not os.path.exists(path):
        os.makedirs(HOST, exist_ok=True)
        cmd = ["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), SOURCES[exe], "-o", path,
               "-L" + os.path.join(ROOT, "rgk_b200"), "-lrgk_b200", "-Wl,-rpath," + os.path.join(ROOT, "rgk_b200")]
        if exe == "rgk_render_multi":
            cmd += ["-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-lcudart", "-lnccl", "-lpthread", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            pytest.fail(f"{path} missing and could not be built: {r.stderr[-500:]}")
    return path


def read_acc(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"RGKACC01"
    w, h, rounds = np.frombuffer(raw, np.uint32, 3, 8)
    s = np.frombuffer(raw, np.float32, int(w) * int(h) * 3, 20).reshape(h, w, 3)
    c = np.frombuffer(raw, np.uint32, int(w) * int(h), 20 + int(w) * int(h) * 12).reshape(h, w)
    return s, c, int(rounds)


def test_pack_roundtrip_python(tmp_path):
    pack, cfg = scenes.material_zoo(width=48, height=32, multisample=4)
    path = str(tmp_path / "zoo.rgkpack")
    pack.save(path, cfg)
    p2, c2 = scene.load_pack(path)
    a, b = pack.arrays(), p2.arrays()
    for k in ("positions", "normals", "tangents", "texcoords", "indices"):
        assert a[k].tobytes() == b[k].tobytes(), k
    assert a["mesh_ranges"] == b["mesh_ranges"] and len(p2.textures) == len(pack.textures)
    assert bytes(pack.desc().sky) == bytes(p2.desc().sky)                    # equal once rounded to the ABI's floats
    for (k1, v1), (k2, v2) in zip(pack.textures, p2.textures):
        assert k1 == k2 and np.array_equal(np.asarray(v1, np.float32), np.asarray(v2, np.float32))
    d1, d2 = pack.desc(), p2.desc()
    assert bytes(d1.materials[2]) == bytes(d2.materials[2]) and d1.n_point_lights == d2.n_point_lights
    assert (c2.xres, c2.yres, c2.multisample, c2.recursion_level) == (cfg.xres, cfg.yres, cfg.multisample, cfg.recursion_level)
    ca, cb = cfg.camera_args(), c2.camera_args()
    assert np.array_equal(np.asarray(ca["pos"], np.float32), cb["pos"]) and np.float32(ca["xview"]) == np.float32(cb["xview"])
    with pytest.raises(ValueError):
        (tmp_path / "bad").write_bytes(b"NOTAPACK" + bytes(64)); scene.load_pack(str(tmp_path / "bad"))


def test_host_layer_against_oracle(tmp_path, oracle):
    exe = _need("host_checks")
    pack, cfg = scenes.load_builtin("cornell-box", width=64, height=48, multisample=4)
    ppath = str(tmp_path / "cornell.rgkpack")
    pack.save(ppath, cfg)
    out = json.loads(subprocess.run([exe, ppath, str(tmp_path)], capture_output=True, text=True, check=True).stdout)
    # EXRTexture arithmetic (src/texture.cpp:334-412), recomputed in fp32
    f = np.float32
    px = np.array([f(0.25) * f(1) + f(2), f(1.5) * f(2), f(0.125) * f(2)], f) + np.array([1, 2, 3], f)
    assert out["count_1_2"] == 7 and np.array_equal(np.array(out["pixel_1_2"], f), px / f(7)) and out["empty"] == [0, 0, 0]
    assert out["normalized_max"] == pytest.approx(1.0, abs=1e-6) and out["scaled_0_0_b"] == float(f(0.125) * f(0.5) / f(2))
    assert out["raw_roundtrip"] is True
    # Imath half(float): round to nearest even, overflow to inf, smallest subnormal
    assert out["half"] == [int(np.float16(x).view(np.uint16)) for x in (1.0, -2.5, 65504.0)] + [0x7c00, 1] + \
        [int(np.float32(0.33333334).astype(np.float16).view(np.uint16)), int((np.float32(1.0009766) + np.float32(0.00048828)).astype(np.float16).view(np.uint16))]
    # the EXR file decodes (OpenCV's OpenEXR codec) to the half-rounded GetPixel values
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    im = cv2.imread(str(tmp_path / "a.exr"), cv2.IMREAD_UNCHANGED)
    if im is not None:          # codec compiled in
        assert im.shape == (3, 4, 4) and np.allclose(im[2, 1, 2::-1], (px / f(7)).astype(np.float16).astype(f)) and np.all(im[..., 3] == 1)
    sraw, craw, rounds = read_acc(str(tmp_path / "a.acc"))
    assert rounds == 7 and craw[2, 1] == 7 and np.array_equal(sraw[2, 1], px)
    # GenerateTaskList == the oracle's restatement of src/render_driver.cpp:30-46
    ot = oracle.generate_tasks(32, 200, 100)
    assert out["tasks"] == [[int(t.x1), int(t.x2), int(t.y1), int(t.y2)] for t in ot]
    # PackFile + Camera: same scene description and camera as the Python side / the oracle
    a = pack.arrays()
    assert out["pack"]["n_vertices"] == 108 and out["pack"]["n_triangles"] == 36 and out["pack"]["n_meshes"] == 9
    assert out["pack"]["possum"] == pytest.approx(float(a["positions"].astype(np.float64).sum()), rel=1e-7)
    assert (out["pack"]["xres"], out["pack"]["multisample"], out["pack"]["depth"]) == (64, 4, 40) and out["pack"]["emission3"] == 17.0
    ca = cfg.camera_args()
    cam = oracle.camera_init(ca["pos"], ca["lookat"], ca["up"], ca["yview"], ca["xview"], ca["xres"], ca["yres"], ca["focus_plane"], ca["lens_size"])
    for k in ("origin", "viewscreen", "viewscreen_x", "viewscreen_y"):
        assert np.array_equal(np.array(out["camera"][k], f), np.array(list(getattr(cam, k)), f)), k
    # host-only commit through the C ABI from C++
    ho = oracle.scene_create(pack.desc())
    oi = oracle.scene_info(ho)
    assert out["host_scene"]["status"] == 0 and out["host_scene"]["n_nodes"] == oi.n_nodes and out["host_scene"]["n_refs"] == oi.n_refs
    assert np.float32(out["host_scene"]["epsilon"]) == np.float32(oi.epsilon)
    # no device in this container: the Scene constructor throws with the library's text (never a CPU fallback)
    import torch
    if not torch.cuda.is_available():
        assert "no CUDA device" in out["scene_ctor"]


@pytest.mark.gpu
def test_cpp_render_frame_matches_python_binding_and_resumes(tmp_path, gpu_ctx):
    """rgk_render (C++: PackFile -> Scene::Commit -> RenderDriver::RenderFrame) == Context.render_frame (Python), bit
    for bit; a run interrupted after 1 round and resumed equals the uninterrupted 3-round run; rendering tile by tile
    through PathTracer::Render (the reference's granularity) gives the same framebuffer."""
    exe = _need("rgk_render")
    pack, cfg = scenes.load_builtin("cornell-box", width=96, height=64, multisample=4)
    ppath = str(tmp_path / "cornell.rgkpack")
    pack.save(ppath, cfg)
    run = lambda *a: json.loads(subprocess.run([exe, ppath] + list(a), capture_output=True, text=True, check=True).stdout)
    st = run(str(tmp_path / "full.exr"), "--rounds", "3", "--raw", str(tmp_path / "full.acc"))
    assert st["rounds"] == 3 and st["samples"] == 3 * 96 * 64 * 4
    s_full, c_full, r_full = read_acc(str(tmp_path / "full.acc"))
    gpu_ctx.commit(pack.desc())
    cam = gpu_ctx.camera(**cfg.camera_args())
    fb, cnt, _ = gpu_ctx.render_frame(cam, cfg.params(), 3)
    assert np.array_equal(cnt, c_full) and fb.tobytes() == s_full.tobytes() and r_full == 3
    # interrupted + resumed
    ck = str(tmp_path / "ck.acc")
    run(str(tmp_path / "part.exr"), "--rounds", "1", "--checkpoint", ck)
    assert read_acc(ck)[2] == 1
    st2 = run(str(tmp_path / "part.exr"), "--rounds", "3", "--checkpoint", ck, "--resume")
    s_res, c_res, r_res = read_acc(ck)
    assert r_res == 3 and st2["rounds"] == 3 and s_res.tobytes() == s_full.tobytes() and np.array_equal(c_res, c_full)
    # tile by tile through PathTracer::Render
    run(str(tmp_path / "tiles.exr"), "--rounds", "3", "--tiles", "--raw", str(tmp_path / "tiles.acc"))
    s_t, c_t, _ = read_acc(str(tmp_path / "tiles.acc"))
    assert np.array_equal(c_t, c_full)
    # per-tile buffers are accumulated in a different order than the one-call round adds rounds: equal up to fp32 sums
    assert np.allclose(s_t, s_full, rtol=1e-6, atol=1e-6)
    # the progressive EXR holds Normalize(output_scale) of the accumulator, as half floats
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    cv2 = pytest.importorskip("cv2")
    im = cv2.imread(str(tmp_path / "full.exr"), cv2.IMREAD_UNCHANGED)
    if im is not None:
        mean = s_full / c_full[..., None]
        want = (mean / mean.max()).astype(np.float32) if cfg.output_scale <= 0 else mean * cfg.output_scale
        assert np.allclose(im[..., 2::-1], want.astype(np.float16).astype(np.float32), atol=2e-3)


@pytest.mark.gpu
def test_native_multi_gpu_driver_on_the_visible_gpus(tmp_path):
    """rgk_render_multi (one thread + context per GPU, ncclReduce per super-round) against the single-GPU driver: same
    counts, same sums.  Uses every visible GPU up to 2 (with one GPU the NCCL communicator has a single rank)."""
    import torch
    if not os.path.exists("/usr/include/nccl.h"):
        pytest.skip("NCCL headers not installed")
    multi, single = _need("rgk_render_multi"), _need("rgk_render")
    n = min(2, torch.cuda.device_count())
    pack, cfg = scenes.load_builtin("cornell-box", width=96, height=64, multisample=4)
    ppath = str(tmp_path / "cornell.rgkpack")
    pack.save(ppath, cfg)
    subprocess.run([single, ppath, str(tmp_path / "s.exr"), "--rounds", "5", "--raw", str(tmp_path / "s.acc")], check=True, capture_output=True)
    out = subprocess.run([multi, ppath, str(tmp_path / "m.exr"), "--gpus", str(n), "--rounds", "5", "--raw", str(tmp_path / "m.acc")],
                         check=True, capture_output=True, text=True).stdout
    st = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert st["gpus"] == n and st["rounds"] == 5 and st["samples"] == 5 * 96 * 64 * 4
    s1, c1, r1 = read_acc(str(tmp_path / "s.acc"))
    s2, c2, r2 = read_acc(str(tmp_path / "m.acc"))
    assert r1 == r2 == 5 and np.array_equal(c1, c2) and np.allclose(s1, s2, rtol=1e-6, atol=1e-6)
