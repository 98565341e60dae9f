"""The other BASELINE configs (stand-in scenes) at reduced resolution / spp: the features they add
(thin-lens camera + envmap sky + sphere light; eight sphere lights in a closed 400 k-triangle hall; a 2 M-triangle
tree 29 levels deep) rendered on the GPU against the CPU oracle with the same sample sequence."""
import numpy as np
import pytest

import raybatches
from rgk_b200 import abi, standin

pytestmark = pytest.mark.gpu


def _render_both(gpu_ctx, oracle, pack, cfg):
    desc = pack.desc()
    gpu_ctx.commit(desc)
    ho = oracle.scene_create(desc)
    assert all(np.array_equal(a, b) for a, b in zip(oracle.scene_kdtree(ho), gpu_ctx.scene_kdtree()))
    cam = gpu_ctx.camera(**cfg.camera_args())
    p = cfg.params(abi.SAMPLER_MT19937)
    tasks = gpu_ctx.generate_tasks(32, p.xres, p.yres)
    fg, cg, sg = gpu_ctx.render_round(cam, p, tasks)
    fo, co, so = oracle.render_round(ho, cam, p, tasks, nthreads=8)
    mean = float(fo.mean())
    rel = abs(float(fg.mean()) - mean) / mean
    rmse = float(np.sqrt(np.mean((fg - fo) ** 2)))
    print(f"{cfg.output_file}: {pack.n_triangles} tris rel_mean={rel:.2e} rmse/mean={rmse / mean:.2e} bit-equal px={np.mean(fg == fo):.3f} "
          f"rays gpu={sg.closest_rays}/{sg.shadow_rays} cpu={so.closest_rays}/{so.shadow_rays}")
    assert np.array_equal(cg, co)
    assert rel <= 2e-3 and rmse <= 0.05 * mean
    assert abs(int(sg.closest_rays) - int(so.closest_rays)) <= 0.002 * so.closest_rays
    return ho


def test_sibenik_standin_lens_envmap(gpu_ctx, oracle):
    pack, cfg = standin.sibenik(width=96, height=54, multisample=4)
    _render_both(gpu_ctx, oracle, pack, cfg)


def test_conference_standin_eight_sphere_lights(gpu_ctx, oracle):
    pack, cfg = standin.conference(width=64, height=36, multisample=4)
    _render_both(gpu_ctx, oracle, pack, cfg)


def test_dragon_sponza_standin_two_million_triangles(gpu_ctx, oracle):
    pack, cfg = standin.dragon_sponza(width=160, height=90, multisample=1)
    cfg.recursion_level = 6
    ho = _render_both(gpu_ctx, oracle, pack, cfg)
    info = gpu_ctx.scene_info()
    assert info.n_triangles > 2_000_000 and info.max_depth >= 28
    cam = gpu_ctx.camera(**cfg.camera_args())
    rays = gpu_ctx.camera_rays(cam, 160, 90, np.stack(np.mgrid[0:90, 0:160][::-1], -1).reshape(-1, 2).astype(np.int32),
                               np.full((160 * 90, 2), 0.5, np.float32))
    hits, sg = gpu_ctx.trace_closest(rays, want_stats=True)
    ho_hits, so = oracle.trace_closest(ho, rays, want_stats=True)
    assert hits.tobytes() == ho_hits.tobytes() and sg.as_dict() == so.as_dict()
    assert sg.prefilter_wrong == 0
