#!/bin/bash
# bench + launch list + one full ncu capture of the mean-shift loop kernel and the FPS kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench exit $?"; cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:meanshift_tc -s 12 -c 2 -o gpurun_out/prof_meanshift -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ms.log 2>&1
echo "ncu meanshift exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_kernel -s 1 -c 1 -o gpurun_out/prof_fps -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fps.log 2>&1
echo "ncu fps exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 3 -o gpurun_out/prof_conv -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv.log 2>&1
echo "ncu conv exit $?"
ls -la gpurun_out
