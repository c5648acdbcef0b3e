#!/bin/bash
# round 2, session G (8 GPUs): multi-GPU parity worker at 8 and 4 ranks (NCCL + peer-memory paths, block8 included),
# weak-scaling bench at N = 1, 2, 4, 8 (default line incl. other_configs), reference arm under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29608 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_8ranks.log 2>&1; echo "worker8 rc=$?"
tail -3 gpurun_out/r02_multi_gpu_worker_8ranks.log
CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 600 $TR --nproc-per-node 4 --master-port 29604 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_4ranks.log 2>&1; echo "worker4 rc=$?"
tail -3 gpurun_out/r02_multi_gpu_worker_4ranks.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench n1 rc=$?"
for N in 2 4 8; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29520+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; echo "bench n$N rc=$?"
done
timeout 300 $TR --nproc-per-node 8 --master-port 29540 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2g_bench_reference_n8.json 2> gpurun_out/r2g_bench_reference_n8.err; echo "ref n8 rc=$?"
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
python - <<'PY'
import json
base=None
for N in (1,2,4,8):
    try:
        d=json.loads(open('gpurun_out/r2g_bench_n%d.json'%N).read().strip().splitlines()[-1])
    except Exception as e:
        print(N,'failed',e); continue
    if N==1: base=d
    oc=d.get('other_configs') or {}
    print(N, round(d['value'],1), round(d['ms_per_step'],2), 'eff', round(d['value']/(N*base['value']),4) if base else None,
          'e2e', round(d['e2e']['value'],1), round(d['e2e']['value']/(N*base['e2e']['value']),4) if base else None, 'host GB/s/rank', round(d['e2e'].get('host_copy_GBs_per_rank',0),2),
          'c5', round(oc.get('c5',{}).get('value',0),1), 'cpu', (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'), 'p2p', d.get('peer_memory_path'))
try:
    d=json.loads(open('gpurun_out/r2g_bench_reference_n8.json').read().strip().splitlines()[-1]); print('ref n8', d['value'], d['cpu_baseline']['cores'], d['n_gpus'])
except Exception as e: print('ref failed', e)
PY
