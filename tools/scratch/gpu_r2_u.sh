#!/bin/bash
# round 2, session U (2 GPUs): 1-D sweeps on segments (ghost values through peer memory): parity worker, C5 / C4 bench at 2 GPUs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 2 --master-port 29616 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_2ranks.log 2>&1; echo "worker2 rc=$?"
tail -14 gpurun_out/r02_multi_gpu_worker_2ranks.log
timeout 300 $TR --nproc-per-node 2 --master-port 29536 bench.py --gpus 2 --config c5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2u_bench_c5_n2.json 2> gpurun_out/r2u_bench_c5_n2.err; echo "bench c5 n2 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2u_bench_c5_n2.json').read().strip().splitlines()[-1]); print('c5 n2', d['value'], d['ms_per_step'], d.get('fuse'))
"
