#!/bin/bash
# the sampler kernels across set sizes and workloads: tools/sampler_workloads.sh [sampler_kernel values, default "0 1"]
run() { echo "== $*"; timeout 300 python bench.py --quick --steps 1 --warmup 3 "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', round(d['ms_per_step'],2), 'ms', {k: round(v,2) for k,v in d['class_ms_per_step'].items()})
    elif 'rror' in l: print(l.strip()[:300])"; }
for k in ${@:-0 1}; do
run --workload sponza --cfg sampler_kernel=$k
run --workload sponza --res 960x540 --spp 256 --cfg sampler_kernel=$k
run --workload sponza --res 480x270 --spp 512 --cfg sampler_kernel=$k
run --workload sponza --res 480x270 --spp 1024 --cfg sampler_kernel=$k
run --workload cornell --cfg sampler_kernel=$k
run --workload sibenik --cfg sampler_kernel=$k
done
