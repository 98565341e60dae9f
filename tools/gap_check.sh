for i in 1 2 3; do
for flag in "" "--no-clocks"; do
python bench.py --quick --steps 6 --warmup 3 $flag 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$flag', round(d['ms_per_step'],2), 'sum', round(sum(d['class_ms_per_step'].values()),2))"
done; done
