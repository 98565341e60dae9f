for v in ${VARIANTS:-2}; do RGK_TRAVERSAL=$v python tools/trace_bench.py --check 100000 $TB_ARGS 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except: print(l.strip()[:300]); continue
    print(d['variant'], d['batch'], d['rays'], round(d['ms'],2), 'ms', round(d['Mrays_s']), 'Mrays/s', round(d['bytes_per_ray']), 'B/ray', round(d['GB_s']), 'GB/s', d.get('bit_exact_vs_oracle'))
"; done
