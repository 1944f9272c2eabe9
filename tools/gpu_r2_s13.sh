#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/batch_ab.py 1 4 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python tools/profile_frame.py 1 4 > gpurun_out/r2_launches.out 2>&1
tail -3 gpurun_out/r2_launches.out
grep -c gpu__time_duration gpurun_out/r2_launches.csv
