#!/usr/bin/env python
"""Host build of the wavefront renderer for the CPU tests: copies rgk_b200/csrc/{*.cuh,*.h,render.cu} into build/host/gen/
with three textual changes that g++ needs -- the product sources are not touched:
  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `doh_launch(kernel, grid, block, smem, stream, args)`  (device_shim.h runs
    the kernel as a loop over blocks and threads);
  * `threadIdx.x & 31` -> `0u`: a warp has one lane in the emulation (ballots return bit 0, shuffles their argument);
  * `extern __shared__ T name[];` -> `static T name[65536];` (k_sampler_mt's tables are private to a lane, so it runs thread by
    thread like the rest; only k_bin, whose threads cooperate through shared memory, is compiled but never run: the tests
    switch the binning off).
With --lanes32 (for device_shim_mt.h, where a CUDA thread is a host thread and a warp has 32 lanes) the lane expressions are left
alone.
usage: gen_host_sources.py <csrc dir> <out dir> [--lanes32]"""
import os, re, sys

src, out = sys.argv[1], sys.argv[2]
lanes32 = "--lanes32" in sys.argv[3:]
os.makedirs(out, exist_ok=True)
launch = re.compile(r'(\bk_\w+(?:<[^<>;]*>)?)\s*<<<(.*?)>>>\s*\(', re.S)     # kernel, optional template arguments, launch configuration
n_launch = 0
for name in sorted(os.listdir(src)):
    if not (name.endswith((".cuh", ".h")) or name == "render.cu"):
        continue
    s = open(os.path.join(src, name)).read()
    s, k = launch.subn(lambda m: "doh_launch(%s, %s, " % (m.group(1), m.group(2)), s)
    n_launch += k
    if not lanes32:
        s = s.replace("(threadIdx.x & 31)", "(0u)").replace("threadIdx.x & 31", "0u")
    s = re.sub(r'extern\s+__shared__\s+(\w+)\s+(\w+)\[\];', r'alignas(16) static \1 \2[65536];', s)      # 256 KB: more than any launch asks for
    dst = os.path.join(out, name if name != "render.cu" else "render_host.inc")
    if not os.path.exists(dst) or open(dst).read() != s:
        open(dst, "w").write(s)
assert n_launch >= 30, n_launch
