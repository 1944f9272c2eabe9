#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "n2 exit $?"; cat gpurun_out/bench_n2.json | python -c "
import json,sys; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2', round(j['value'],1), round(j['e2e']['value'],1), 'serial', round(j['serial']['value'],1), j['config']['multi_gpu'])"
tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
echo "ref n2 exit $?"; cat gpurun_out/bench_ref_n2.json
