#!/bin/bash
# final validation of the round: all GPU tests, smoke, bench, ncu launch list, ncu --set full of the two cooperative kernels
mkdir -p gpurun_out
bash tools/gpu_full.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:meanshift_tc_persistent -c 1 -o gpurun_out/prof_loop_persistent -f python tools/one_frame.py > gpurun_out/ncu_loop.log 2>&1; echo "ncu loop exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps4_kernel -c 1 -o gpurun_out/prof_fps_tc -f python tools/one_frame.py > gpurun_out/ncu_fps.log 2>&1; echo "ncu fps exit $?"
ls -la gpurun_out/*.ncu-rep
