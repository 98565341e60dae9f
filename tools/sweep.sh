run() { echo "== $*"; env "$@" python tools/trace_bench.py --reps 3 --check 60000 2>&1 | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:200]); continue
    out.append('%s %d%s' % (d['batch'], round(d['Mrays_s']), '' if d.get('bit_exact_vs_oracle', True) else ' MISMATCH'))
print('   ', ' | '.join(out))
"; }
run RGK_TRAVERSAL=6 RGK_REFILL=8
for si in 3 6; do for sl in 4 8 16; do run RGK_TRAVERSAL=4 RGK_REFILL=8 RGK_STEPS_INNER=$si RGK_STEPS_LEAF=$sl; done; done
