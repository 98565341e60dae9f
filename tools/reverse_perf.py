import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np
from rgk_b200 import device, scenes, standin
ctx = device.Context(0)
for name, (pack, cfg), rev, depth in (("cornell 256x256x16", scenes.load_builtin("cornell-box", width=256, height=256, multisample=16), 3, 10),
                                        ("sponza-standin 480x270x16", standin.sponza(width=480, height=270, multisample=16), 2, 4)):
    ctx.commit(pack.desc())
    cam = ctx.camera(**cfg.camera_args())
    p = cfg.params(); p.depth = depth
    tasks = ctx.generate_tasks(32, p.xres, p.yres)
    for r in (0, rev):
        p.reverse = r
        ctx.render_round(cam, p, tasks)
        t = time.time(); fb, cnt, st = ctx.render_round(cam, p, tasks); dt = time.time() - t
        print(name, "reverse", r, "ms", round(st.gpu_ms, 1), "closest", int(st.closest_rays), "shadow", int(st.shadow_rays), "mean", float(fb.mean() / p.multisample), "finite", bool(np.isfinite(fb).all()))
