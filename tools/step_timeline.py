#!/usr/bin/env python
"""Where does a step's wall time go?  Per step: wall time of the rgk_render_round_device call, the device time between the call's
own first and last event (rgk_round_stats.gpu_ms), the sum of its kernel-class events, and the wall time between calls."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rgk_b200 import abi, device, multigpu
pack, cfg, label = bench.build_workload("sponza", None, None)
ctx = device.Context(0, stream=torch.cuda.current_stream().cuda_stream)
ctx.commit(pack.desc())
cam = ctx.camera(**cfg.camera_args())
p = cfg.params(abi.SAMPLER_MT19937)
tasks = ctx.generate_tasks(32, p.xres, p.yres)
total = torch.zeros((p.yres, p.xres, 3), dtype=torch.float32, device="cuda")
cnt = torch.zeros((p.yres, p.xres), dtype=torch.int32, device="cuda")
last = None
for i in range(10):
    t0 = time.perf_counter()
    st = ctx.render_round_device(cam, p, tasks, total.data_ptr(), cnt.data_ptr(), 42, i * len(tasks))
    t1 = time.perf_counter()
    cls = float(st.closest_ms) + float(st.shadow_ms) + float(st.shade_ms) + float(st.sampler_ms)
    print("step %d: call %.2f ms  gpu(ev0..ev1) %.2f  classes %.2f  since last call %.2f" % (i, (t1 - t0) * 1e3, float(st.gpu_ms), cls, (t0 - last) * 1e3 if last else 0.0), flush=True)
    last = t1
