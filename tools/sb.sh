for lib in "" inc9 inc10; do echo "lib=$lib"; if [ -n "$lib" ]; then export RGK_B200_LIB=$PWD/rgk_b200/librgk_b200_$lib.so; else unset RGK_B200_LIB; fi; python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','samples_per_s')}); print(d['kernel_share_of_step'])"; done
