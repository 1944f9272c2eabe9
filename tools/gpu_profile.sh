#!/bin/bash
# ncu evidence of a build: launch list of one batch-1 frame and one batch-4 step, full captures of the top kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_frame.py 1 4 > gpurun_out/r2_launches.out 2>&1
grep -c gpu__time_duration gpurun_out/r2_launches.csv
for k in conv_pair_kernel conv_wres_kernel fps4_kernel meanshift_tc_persistent_kernel head_row_kernel; do
  # profile_frame.py runs 4 passes per batch size (eager warm-up, then 3 graph replays of the backbone); the LAST launches are
  # the batch-4 step: skip the earlier passes of every kernel where possible
  case $k in
    conv_pair_kernel) skip=46; cnt=2 ;;       # 4 x 6 (batch 1) + 19 (batch 4 warm-up) + 3: two layer-3 launches of a 4-frame step
    conv_wres_kernel) skip=52; cnt=14 ;;      # 4 x 13 (batch 1): the 6 layer-1 + 7 layer-2 launches of a 4-frame step (+ 1)
    fps4_kernel) skip=3; cnt=1 ;;             # the first batch-4 field
    meanshift_tc_persistent_kernel) skip=2; cnt=2 ;;   # last batch-1 launch, first batch-4 launch
    head_row_kernel) skip=5; cnt=1 ;;
  esac
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o gpurun_out/r2_prof_$k -f python tools/profile_frame.py 1 4 > gpurun_out/r2_prof_$k.log 2>&1
  tail -2 gpurun_out/r2_prof_$k.log
done
ls -la gpurun_out/*.ncu-rep
