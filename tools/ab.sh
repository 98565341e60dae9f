#!/bin/bash
# A/B of environment-selected variants on the bench workload: tools/ab.sh "VAR=1 VAR2=x" "VAR=0" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('  ', round(d['ms_per_step'],1), 'ms', round(d['value']), 'Mrays/s', {k: round(v*d['ms_per_step'],1) for k,v in d['kernel_share_of_step'].items()})"
done
