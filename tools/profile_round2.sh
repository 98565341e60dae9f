#!/bin/bash
# ncu inventory of round 2 (B200_PROFILING.md recipe), every command first run without ncu:
#  (1) launch list (gpu__time_duration per launch) of the headline bench command at full size;
#  (2) `--set full` capture of one whole round of the headline pipeline at 64 spp, 640x360 (ncu's per-pass save / restore of
#      the multi-GB path state makes a full-size capture of all 17 launches impractical; queue shapes and per-ray / per-vertex /
#      per-pixel figures are those of the headline);
#  (3) `--set full` capture of the dominant kernels at FULL size (k_shade first bounce, k_closest_bvh bounce launch, k_sampler_warp);
#  (4) `--set full` capture of the bidirectional mode's kernels (tools/reverse_perf.py).
#   gpurun --timeout 2400 -- 'bash tools/profile_round2.sh'
mkdir -p gpurun_out
T="timeout 600"
CMD="python bench.py --quick --steps 1 --warmup 3"
$T $CMD > gpurun_out/r2f_plain_full.json 2> gpurun_out/r2f_plain_full.err || exit 1
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_full.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1
CMDS="$CMD --res 640x360"
$T $CMDS > gpurun_out/r2f_plain_360.json 2> gpurun_out/r2f_plain_360.err || exit 1
# one whole round: launches 52..68 of the k_* kernels (3 warm-up rounds of 17 launches each come first)
$T ncu --set full --clock-control none --import-source on -k regex:"^k_|::k_" -s 51 -c 17 -o gpurun_out/r2f_round_360 -f $CMDS > gpurun_out/r2f_ncu_round.log 2>&1
tail -n 2 gpurun_out/r2f_ncu_round.log
# full size: the timed round's k_sampler_warp, first k_closest_bvh + k_shade<0>, second k_closest_bvh + k_shade<1>, both k_shadow_bvh
$T ncu --set full --clock-control none --import-source on -k regex:"k_sampler_warp|k_closest_bvh|k_shade|k_shadow_bvh" -s 21 -c 7 -o gpurun_out/r2f_top_full -f $CMD > gpurun_out/r2f_ncu_top.log 2>&1
tail -n 2 gpurun_out/r2f_ncu_top.log
$T python tools/reverse_perf.py > gpurun_out/r2f_reverse_plain.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:"k_shade_rev|k_shadow_rev|k_lightgen|k_connect|k_assemble" -c 12 -o gpurun_out/r2f_reverse -f python tools/reverse_perf.py > gpurun_out/r2f_ncu_reverse.log 2>&1
tail -n 2 gpurun_out/r2f_ncu_reverse.log
