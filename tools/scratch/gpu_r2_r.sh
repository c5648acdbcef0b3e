#!/bin/bash
# round 2, session R (4 GPUs): one-sweep GMRES on slabs at 4 ranks (distinct up / down neighbours): parity worker, bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 4 --master-port 29614 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_4ranks.log 2>&1; echo "worker4 rc=$?"
tail -6 gpurun_out/r02_multi_gpu_worker_4ranks.log
timeout 300 $TR --nproc-per-node 4 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2r_bench_n4.json 2> gpurun_out/r2r_bench_n4.err; echo "bench n4 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2r_bench_n4.json').read().strip().splitlines()[-1]); print('n4', d['value'], d['ms_per_step'], d.get('fuse'), (d.get('e2e') or {}).get('value'), d['roofline']['frac'], d['roofline'].get('frac_step'))
print({k:(v.get('value'), v.get('ms_per_step')) for k,v in (d.get('other_configs') or {}).items()} if isinstance(d.get('other_configs'), dict) else d.get('other_configs'))
"
