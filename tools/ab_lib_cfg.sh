run() { echo "== $1 | $2"; RGK_B200_LIB=$1 timeout 120 python bench.py --quick --steps 3 --warmup 3 --cfg "$2" 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  ', round(d['ms_per_step'],2), 'ms', round(d['value']), 'Mrays/s', {k: round(v,2) for k,v in d['class_ms_per_step'].items()})
    elif 'rror' in l: print(l.strip()[:300])"; }
# usage: tools/ab_lib_cfg.sh "<lib suffix or empty>|<cfg>" ...   e.g.  "m4|sampler_slots=3" "|sampler_slots=6"
for a in "$@"; do sfx="${a%%|*}"; cfg="${a#*|}"; run rgk_b200/librgk_b200${sfx:+_$sfx}.so "$cfg"; done
