#!/bin/bash
# A/B of rgk_device_cfg settings on the bench workload: tools/ab.sh "binning=0" "refill_shadow=20,arb_grid=4" "" ...
for cfg in "$@"; do
  echo "== $cfg"
  python bench.py --steps 2 --warmup 3 --no-cpu --cfg "$cfg" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('  ', round(d['ms_per_step'],1), 'ms', round(d['value']), 'Mrays/s', {k: round(v*d['ms_per_step'],1) for k,v in d['kernel_share_of_step'].items()})"
done
