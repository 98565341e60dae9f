for th in 64 128 256; do for rf in 4 8 16 32; do
echo "threads=$th refill=$rf"; RGK_TRACE_THREADS=$th RGK_REFILL=$rf python tools/trace_bench.py --reps 3 2>&1 | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except: continue
    out.append('%s %d' % (d['batch'], round(d['Mrays_s'])))
print('   ', ' | '.join(out))
"; done; done
