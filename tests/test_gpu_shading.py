"""Unit-level parity of the device shading functions against the CPU oracle through rgk_probe, plus an image-level
check on the material-zoo scene (every live BxDF, image textures, bump maps, lens, sphere + areal light, envmap).

Tolerance: these functions call sinf/cosf/acosf/asinf/atan2f/sqrtf; CUDA's implementations differ from glibc's by
<= 2 ulp, which the LTC matrix inverse amplifies mildly.  Stated bound: 1e-5 relative (2e-4 for LTC lobes whose
3x3 inverses are ill-conditioned near grazing angles), absolute floor 1e-6."""
import os

import numpy as np
import pytest

from rgk_b200 import abi, scenes

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, rel, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-1)
    bad = ~(np.isclose(a, b, rtol=rel, atol=1e-6) | (np.isnan(a) & np.isnan(b)))
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} outside {rel}: worst {err.max():.3e}"


@pytest.fixture(scope="module")
def zoo(gpu_ctx, oracle):
    pack, cfg = scenes.material_zoo(width=48, height=32, multisample=4, lens=0.05)
    desc = pack.desc()
    gpu_ctx.commit(desc)
    return pack, cfg, desc, oracle.scene_create(desc), np.load(os.path.join(G, "zoo.npz"))


def test_bxdf_value_and_sample_all_kinds(gpu_ctx, oracle, zoo):
    pack, cfg, desc, ho, g = zoo
    Vi, Vr, uv, smp = g["Vi"], g["Vr"], g["uv"], g["smp"]
    for mi, m in enumerate(pack.materials):
        ltc = m["bxdf"] >= abi.BXDF_LTC_BECKMANN or m["bxdf"] == abi.BXDF_MIX
        rel = 2e-4 if ltc else 1e-5
        val = gpu_ctx.probe(abi.PROBE_BXDF_VALUE, mi, np.concatenate([Vi, Vr, uv], 1))
        _close(val, oracle.bxdf_value(ho, mi, Vi, Vr, uv), rel, f"value[{m['name']}]")
        _close(val, g[f"value_{mi}"], rel, f"value[{m['name']}] vs reference fixture")
        s = gpu_ctx.probe(abi.PROBE_BXDF_SAMPLE, mi, np.concatenate([Vi, uv, smp], 1))
        so = oracle.bxdf_sample(ho, mi, Vi, uv, smp)
        assert np.array_equal(s[:, 6], so[:, 6])                      # may_leak flag is discrete: exact
        _close(s[:, :6], so[:, :6], rel, f"sample[{m['name']}]")


def test_textures_lights_sky_frames(gpu_ctx, oracle, zoo):
    pack, cfg, desc, ho, g = zoo
    for ti in range(desc.n_textures):
        got = gpu_ctx.probe(abi.PROBE_TEXTURE, ti, g["tuv"])
        assert np.array_equal(got.view(np.uint32), g[f"tex_{ti}"].view(np.uint32)), ti   # no transcendental: bit-exact
    lights = gpu_ctx.probe(abi.PROBE_RANDOM_LIGHT, 0, g["light_samples"])
    ref = oracle.random_light(ho, g["light_samples"])
    assert np.array_equal(lights.view(np.uint32), ref.view(np.uint32))                    # pure arithmetic: bit-exact
    _close(gpu_ctx.probe(abi.PROBE_SKY, 0, g["sky_dirs"]), g["sky"], 1e-4, "envmap sky")
    rng = np.random.default_rng(9)
    nrm = rng.normal(size=(512, 3)).astype(np.float32)
    nrm[:4] = [[0, 0, -1], [1e-4, 0, -1], [0, 0, 1], [0, 1e-3, -1]]                     # antipodal special case of RotationBetweenVectors
    v = rng.normal(size=(512, 3)).astype(np.float32)
    fr = gpu_ctx.probe(abi.PROBE_FRAME, 0, np.concatenate([nrm, v], 1))
    _close(fr[:, 3:], v, 1e-4, "toGlobal(toLocal(v)) == v")
    n_unit = nrm / np.linalg.norm(nrm, axis=1, keepdims=True)
    fz = gpu_ctx.probe(abi.PROBE_FRAME, 0, np.concatenate([nrm, n_unit], 1))
    # rows 0,1,3 take the antipodal branch (src/glm.cpp:10-21): a half-turn about an axis perpendicular to the normal,
    # which maps n to -n, i.e. to +Z only up to the normal's own deviation from -Z -- as in the reference
    anti = n_unit[:, 2] < -1 + 0.001
    assert anti.sum() >= 3
    assert np.abs(fz[~anti, :3] - [0, 0, 1]).max() < 1e-5     # toLocal(normal) == +Z
    assert np.abs(fz[anti, :3] - [0, 0, 1]).max() < 0.05      # -n is within asin(sqrt(0.002)) of +Z


def test_zoo_image_same_sample_sequence(gpu_ctx, oracle, zoo):
    """All BxDFs in one image, same sampler sequence as the CPU: rel-mean <= 2e-3, RMSE <= 5% of the mean
    (mirror / dielectric paths amplify last-ulp differences of the sampled directions into different hits)."""
    pack, cfg, desc, ho, g = zoo
    pack2, cfg2 = scenes.material_zoo(width=96, height=64, multisample=16, lens=0.05)
    desc2 = pack2.desc()
    gpu_ctx.commit(desc2)
    ho2 = oracle.scene_create(desc2)
    cam = gpu_ctx.camera(**cfg2.camera_args())
    p = cfg2.params(abi.SAMPLER_MT19937)
    tasks = gpu_ctx.generate_tasks(32, p.xres, p.yres)
    fg, cg, sg = gpu_ctx.render_round(cam, p, tasks)
    fo, co, so = oracle.render_round(ho2, cam, p, tasks)
    mean = float(fo.mean())
    rel_mean = abs(float(fg.mean()) - mean) / mean
    rmse = float(np.sqrt(np.mean((fg - fo) ** 2)))
    print(f"zoo 96x64x16: rel_mean={rel_mean:.3e} rmse/mean={rmse / mean:.3e} bit-equal pixels={np.mean(fg == fo):.3f} "
          f"rays gpu={sg.closest_rays}/{sg.shadow_rays} cpu={so.closest_rays}/{so.shadow_rays}")
    assert np.array_equal(cg, co)
    assert rel_mean <= 2e-3 and rmse <= 0.05 * mean
    assert abs(int(sg.closest_rays) - int(so.closest_rays)) <= 0.002 * so.closest_rays
    gpu_ctx.commit(desc)   # restore the module fixture's scene
