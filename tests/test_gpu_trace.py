"""Parity of the CUDA traversal against the CPU oracle, through the C ABI (rgk_trace_closest /
rgk_trace_shadow).  Bit-exact: triangle index, hit flag, t and barycentrics (integer / IEEE work)."""
import numpy as np
import pytest

import raybatches

pytestmark = pytest.mark.gpu


def _setup(gpu_ctx, oracle, cornell, use_oracle_tree):
    pack, cfg, desc = cornell
    ho = oracle.scene_create(desc)
    tree = None
    if use_oracle_tree:
        from rgk_b200 import abi
        nodes, refs = oracle.scene_kdtree(ho)
        tree = abi.KdTree(len(nodes) // 2, nodes.ctypes.data_as(abi.u32p), len(refs), refs.ctypes.data_as(abi.u32p))
        tree._keep = (nodes, refs)
    gpu_ctx.commit(desc, tree)
    cam = gpu_ctx.camera(**cfg.camera_args())
    return ho, cam


@pytest.mark.parametrize("use_oracle_tree", [True, False])
def test_cornell_tree_and_primary_hits(gpu_ctx, oracle, cornell, use_oracle_tree):
    ho, cam = _setup(gpu_ctx, oracle, cornell, use_oracle_tree)
    # the host-built tree equals the reference-procedure tree word for word
    no, ro = oracle.scene_kdtree(ho)
    ng, rg = gpu_ctx.scene_kdtree()
    assert np.array_equal(no, ng) and np.array_equal(ro, rg)
    io, ig = oracle.scene_info(ho), gpu_ctx.scene_info()
    assert io.epsilon == ig.epsilon and list(io.bbox) == list(ig.bbox) and io.max_depth == ig.max_depth
    rays = raybatches.primary(oracle, oracle.camera_init(**_cam_kw(cornell)), 256, 256)
    rays_gpu = gpu_ctx.camera_rays(cam, 256, 256, _grid(256, 256), np.full((65536, 2), 0.5, np.float32))
    assert rays.tobytes() == rays_gpu.tobytes()          # raygen kernel == Camera::GetPixelRay
    hg, sg = gpu_ctx.trace_closest(rays, want_stats=True)
    ho_hits, so = oracle.trace_closest(ho, rays, want_stats=True)
    assert hg.tobytes() == ho_hits.tobytes()
    assert sg.as_dict() == so.as_dict()                  # same nodes / refs / tests visited: same algorithm
    assert sg.prefilter_wrong == 0 and sg.prefiltered > 0   # the 2-D bounds pre-filter only rejects what the exact test rejects
    assert (hg["triangle"] != 0xFFFFFFFF).mean() > 0.99


def _cam_kw(cornell):
    ca = cornell[1].camera_args()
    return dict(pos=ca["pos"], lookat=ca["lookat"], up=ca["up"], yview=ca["yview"], xview=ca["xview"], xres=ca["xres"],
                yres=ca["yres"], focus_plane=ca["focus_plane"], lens_size=ca["lens_size"])


def _grid(w, h):
    ys, xs = np.mgrid[0:h, 0:w]
    return np.stack([xs.ravel(), ys.ravel()], 1).astype(np.int32)


def test_cornell_bounce_and_shadow(gpu_ctx, oracle, cornell):
    ho, cam = _setup(gpu_ctx, oracle, cornell, False)
    rays = raybatches.primary(oracle, oracle.camera_init(**_cam_kw(cornell)), 256, 256, jitter_seed=3)
    hits = oracle.trace_closest(ho, rays)
    planes = oracle.scene_planes(ho)
    eps = oracle.scene_info(ho).epsilon
    brays, ign = raybatches.bounce(rays, hits, planes[:, :3], eps)
    hg = gpu_ctx.trace_closest(brays, ign)
    hc = oracle.trace_closest(ho, brays, ign)
    assert hg.tobytes() == hc.tobytes()
    assert not np.any(hg["triangle"] == ign)             # the ignored triangle is never returned
    # without the ignore list some rays re-hit their source triangle: a different code path, also exact
    assert gpu_ctx.trace_closest(brays).tobytes() == oracle.trace_closest(ho, brays).tobytes()
    a, b = raybatches.shadow_segments(rays, hits, (-0.005, 1.97, -0.03))
    vg, sg = gpu_ctx.trace_shadow(a, b, want_stats=True)
    vc = oracle.trace_shadow(ho, a, b)
    assert np.array_equal(vg, vc) and sg.prefilter_wrong == 0
    assert 0.05 < vg.mean() < 0.95


def test_edge_cases(gpu_ctx, oracle, cornell):
    ho, cam = _setup(gpu_ctx, oracle, cornell, False)
    from checkers import RAY_DT
    # empty batch
    assert len(gpu_ctx.trace_closest(np.zeros(0, RAY_DT))) == 0
    # rays missing the scene box, axis-parallel rays (1/0 = inf in the slab test), rays starting on a split
    # plane, zero direction (NaN everywhere), ragged count (not a multiple of 32)
    r = np.zeros(77, RAY_DT)
    r["tfar"] = 10000.0
    rng = np.random.default_rng(5)
    r["origin"] = rng.uniform(-3, 3, (77, 3)).astype(np.float32)
    d = rng.normal(size=(77, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["direction"] = d
    r["direction"][:8] = [0, 0, -1]
    r["origin"][:8, :2] = rng.uniform(-0.9, 0.9, (8, 2)); r["origin"][:8, 1] += 1; r["origin"][:8, 2] = 5
    r["direction"][8:12] = [1, 0, 0]
    r["origin"][12] = [0, 1, 0]; r["direction"][12] = [0, 1, 0]
    r["direction"][13] = [0, 0, 0]
    r["origin"][14] = [0, 5, 0]; r["direction"][14] = [0, 1, 0]       # pointing away from the box
    r["tfar"][15] = 0.5                                                # far plane before any surface
    hg = gpu_ctx.trace_closest(r)
    hc = oracle.trace_closest(ho, r)
    assert hg.tobytes() == hc.tobytes()
