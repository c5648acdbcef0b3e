#!/bin/bash
# round 2, session P (2 GPUs): one-sweep GMRES on slabs (ghost rows of every basis vector through peer memory): parity
# worker, weak-scaling bench at 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 2 --master-port 29612 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_2ranks.log 2>&1; echo "worker2 rc=$?"
tail -12 gpurun_out/r02_multi_gpu_worker_2ranks.log
timeout 400 $TR --nproc-per-node 2 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench_n2.json 2> gpurun_out/r2p_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2p_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r2p_bench_n2.json').read().strip().splitlines()[-1]); print('n2', d['value'], d['ms_per_step'], d.get('fuse'), (d.get('e2e') or {}).get('value'), d['roofline']['frac'], d['roofline'].get('frac_step'))
"
