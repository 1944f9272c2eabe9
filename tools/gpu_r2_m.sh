#!/bin/bash
# session-2 call M: ncu --set full of a tiny conv launch (1x1 downsample, 4 K blocks) and a layer-4 3x3 launch: where do 8 us go?
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:conv_tc_kernel --launch-skip 43 --launch-count 2 -o gpurun_out/prof_conv_small -f python tools/one_frame.py > gpurun_out/ncu_conv_small.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:conv_tc_kernel --launch-skip 67 --launch-count 1 -o gpurun_out/prof_conv_l4 -f python tools/one_frame.py > gpurun_out/ncu_conv_l4.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/*.ncu-rep
