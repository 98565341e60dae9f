run() { echo "== $*"; timeout 300 python bench.py --quick --steps 1 --warmup 3 "$@" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', round(d['ms_per_step'],2), 'ms', {k: round(v,2) for k,v in d['class_ms_per_step'].items()})
    elif 'rror' in l: print(l.strip()[:300])"; }
for k in 0 1; do
for spp in 1 4 16 36; do run --workload sponza --spp $spp --cfg sampler_kernel=$k; done
done
