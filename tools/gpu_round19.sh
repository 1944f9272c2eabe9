#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
UOC_CONV_HALO=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 78 -c 2 -o gpurun_out/prof_halo2 -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --depth 1 > gpurun_out/ncu_halo.log 2>&1
echo "ncu exit $?"
