for sm in 1 0; do echo "RGK_SAMPLER_SMEM=$sm"; RGK_SAMPLER_SMEM=$sm python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','samples_per_s')}); print(d['kernel_share_of_step'])"; done
