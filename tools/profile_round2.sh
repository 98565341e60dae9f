#!/bin/bash
# ncu inventory of round 2 (B200_PROFILING.md recipe), every command first run without ncu:
#  (1) launch list (gpu__time_duration per launch) of the headline bench command at full size;
#  (2) `--set full` capture of one whole round of the headline pipeline at 64 spp, 640x360 (ncu's per-pass save / restore of
#      the multi-GB path state makes a full-size capture of all 17 launches impractical; queue shapes and per-ray / per-vertex /
#      per-pixel figures are those of the headline);
#  (3) `--set full` capture of the dominant kernels at FULL size (k_sampler_warp, both k_closest_bvh, both k_shade, both k_shadow_bvh);
#  (4) `--set full` capture of the bidirectional mode's kernels (tools/reverse_perf.py).
# gpurun brings back at most 64 MiB: each report is turned into its CSV / text views on the box and then deleted.
#   gpurun --timeout 2400 -- 'bash tools/profile_round2.sh'
mkdir -p gpurun_out
O=gpurun_out
T="timeout 600"
views() {   # views <report stem> <launch indices for the source-line / basic-block views>
  ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv 2>/dev/null
  python tools/ncu_keys.py $O/$1.ncu-rep > $O/$1_keys.txt 2>&1
  shift_stem=$1; shift
  for i in "$@"; do
    python tools/ncu_lines.py $O/$shift_stem.ncu-rep $i 60 > $O/${shift_stem}_lines_$i.txt 2>&1
    python tools/ncu_segments.py $O/$shift_stem.ncu-rep $i 30 > $O/${shift_stem}_blocks_$i.txt 2>&1
  done
  rm -f $O/$shift_stem.ncu-rep
}
CMD="python bench.py --quick --steps 1 --warmup 3"
$T $CMD > $O/r2f_plain_full.json 2> $O/r2f_plain_full.err || exit 1
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_full.csv $CMD > $O/r2f_ncu_launches.log 2>&1
CMDS="$CMD --res 640x360"
$T $CMDS > $O/r2f_plain_360.json 2> $O/r2f_plain_360.err || exit 1
# one whole round: launches 52..68 of the k_* kernels (3 warm-up rounds of 17 launches each come first)
$T ncu --set full --clock-control none --import-source on -k regex:"^k_|::k_" -s 51 -c 17 -o $O/r2f_round_360 -f $CMDS > $O/r2f_ncu_round.log 2>&1
tail -n 1 $O/r2f_ncu_round.log
python tools/make_kernel_metrics.py $O/r2f_round_360.ncu-rep $O/r2f_plain_360.json $O/r2f_kernel_metrics.json $O/r2f_kernel_metrics.md > $O/r2f_kernel_metrics.log 2>&1
views r2f_round_360 2 5
# full size: the timed round's k_sampler_warp, k_closest_bvh + k_shade<0> + k_shadow_bvh of the camera rays, the same of the bounce
$T ncu --set full --clock-control none --import-source on -k regex:"k_sampler_warp|k_closest_bvh|k_shade|k_shadow_bvh" -s 21 -c 7 -o $O/r2f_top_full -f $CMD > $O/r2f_ncu_top.log 2>&1
tail -n 1 $O/r2f_ncu_top.log
views r2f_top_full 0 1 2 3 4 5 6
$T python tools/reverse_perf.py > $O/r2f_reverse_plain.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:"k_shade_rev|k_shadow_rev|k_lightgen|k_connect|k_assemble" -c 12 -o $O/r2f_reverse -f python tools/reverse_perf.py > $O/r2f_ncu_reverse.log 2>&1
tail -n 1 $O/r2f_ncu_reverse.log
views r2f_reverse
du -sh $O
