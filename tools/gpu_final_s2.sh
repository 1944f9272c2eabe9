#!/bin/bash
# session-2 final validation: all GPU tests in one process (the driver's command), smoke, bench, configs 3 / 5,
# ncu launch list of the bench command, ncu --set full of the persistent loop kernel
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; tail -1 gpurun_out/build.log
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/t_all_gpu.log 2>&1; echo "gpu tests exit $?"; tail -2 gpurun_out/t_all_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 40 --warmup 3 > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; echo "bench exit $?"
python -c "
import json; j=json.load(open('gpurun_out/bench_s2.json')); print(round(j['value'],1), round(j['e2e']['value'],1), round(j['e2e_raw_inputs']['value'],1), 'serial', round(j['serial']['value'],1), j['stages_ms'], round(j['roofline']['frac'],4), j['clocks'], j['cpu_baseline']['value'], j['gpu_launches'])"
timeout 300 python tools/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "configs exit $?"; cut -c1-200 gpurun_out/bench_configs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_s2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:meanshift_tc_persistent -c 1 -o gpurun_out/prof_loop_s2 -f python tools/one_frame.py > gpurun_out/ncu_loop.log 2>&1; echo "ncu loop exit $?"
ls -la gpurun_out/*.ncu-rep
