#!/bin/bash
# step time against the sum of its kernel classes, several processes (host jitter shows as a gap that changes from run to run)
for i in 1 2 3 4; do
python bench.py --quick --steps 6 --warmup 3 "$@" 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(round(d['ms_per_step'],2), 'sum', round(sum(d['class_ms_per_step'].values()),2), {k: round(v,2) for k,v in d['class_ms_per_step'].items()})"
done
