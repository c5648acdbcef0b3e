#!/bin/bash
# round 2, session Y (8 GPUs): the default bench line at 8 ranks (one-sweep GMRES on 8 slabs, DG on 8 segments)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 8 --master-port 29538 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r2y_bench_n8.err | cut -c1-300
python -c "
import json
d=json.loads(open('gpurun_out/r2y_bench_n8.json').read().strip().splitlines()[-1]); print('n8', d['value'], d['ms_per_step'], d.get('fuse'), (d.get('e2e') or {}).get('value'), d['roofline']['frac'])
print({k:(v.get('value'), v.get('ms_per_step')) for k,v in (d.get('other_configs') or {}).items() if isinstance(v, dict)})
"
