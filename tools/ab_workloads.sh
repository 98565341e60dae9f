for w in cornell sibenik; do for b in 1 0; do echo "== $w RGK_BIN=$b"; RGK_BIN=$b python bench.py --workload $w --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('  ', round(d['ms_per_step'],1), 'ms', round(d['value']), 'Mrays/s', round(d['samples_per_s']/1e6,1), 'Msamples/s', d['gpu_launches'], {k: round(v*d['ms_per_step'],1) for k,v in d['kernel_share_of_step'].items()})"; done; done
