#!/bin/bash
# ncu inventory of one round of the headline pipeline (B200_PROFILING.md recipe).  64 spp keeps the sampler / queue shapes of
# the headline; the resolution is cut to 640x360 so that ncu's per-pass save / restore of the path state stays short.
#   gpurun --timeout 1500 -- 'bash tools/profile_round2.sh'
CMD="python bench.py --quick --steps 1 --warmup 3 --res 640x360"
mkdir -p gpurun_out
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
# one whole round: launches 52..68 of the k_* kernels (3 warm-up rounds of 17 launches each come first)
timeout 1100 ncu --set full --clock-control none --import-source on -k regex:"^k_|::k_" -s 51 -c 17 -o gpurun_out/r2_prof_render -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
tail -n 3 gpurun_out/r2_ncu_full.log
