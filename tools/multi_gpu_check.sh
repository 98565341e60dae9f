set -e
python - <<'PY'
import sys; sys.path.insert(0, '.')
from rgk_b200 import scenes
p, c = scenes.load_builtin('cornell-box', width=128, height=96, multisample=4)
p.save('gpurun_out/cornell_multi.rgkpack', c)
PY
build/host/rgk_render gpurun_out/cornell_multi.rgkpack gpurun_out/m1.exr --rounds 6 --raw gpurun_out/m1.acc
build/host/rgk_render_multi gpurun_out/cornell_multi.rgkpack gpurun_out/m2.exr --gpus 2 --rounds 6 --raw gpurun_out/m2.acc
build/host/rgk_render_multi gpurun_out/cornell_multi.rgkpack gpurun_out/m3.exr --gpus 2 --rounds 5 --raw gpurun_out/m3.acc
python - <<'PY'
import numpy as np
def acc(path):
    raw = open(path, 'rb').read(); w, h, r = np.frombuffer(raw, np.uint32, 3, 8)
    s = np.frombuffer(raw, np.float32, int(w)*int(h)*3, 20); c = np.frombuffer(raw, np.uint32, int(w)*int(h), 20 + int(w)*int(h)*12)
    return s, c, int(r)
s1, c1, r1 = acc('gpurun_out/m1.acc'); s2, c2, r2 = acc('gpurun_out/m2.acc'); s3, c3, r3 = acc('gpurun_out/m3.acc')
print('counts equal', np.array_equal(c1, c2), 'rounds', r1, r2, 'max rel diff', float(np.abs(s1 - s2).max() / s1.max()), 'allclose', np.allclose(s1, s2, rtol=1e-5, atol=1e-5))
print('5 rounds on 2 GPUs: count', int(c3.min()), int(c3.max()))
PY
# a GPU whose round fails must not hang the others in the collective: the run ends with an error status
set +e
timeout 60 build/host/rgk_render_multi gpurun_out/cornell_multi.rgkpack gpurun_out/m4.exr --gpus 2 --rounds 6 --fail-gpu 1
echo "fail-gpu run: exit status $? (1 = reported, 124 = hung)"
