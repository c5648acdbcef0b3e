#!/bin/bash
# round 2, session E (2 GPUs): multi-GPU parity worker (NCCL + peer-memory paths), then 1-GPU tests + bench, 2-GPU bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r2e_multi2.log 2>&1
echo "multi rc=$?" >> gpurun_out/r2e_multi2.log
tail -30 gpurun_out/r2e_multi2.log
timeout 600 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "bench n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "bench n2 rc=$?"
tail -c 300 gpurun_out/r2e_bench_n2.err
python -c "
import json
for f in ('r2e_bench_n1','r2e_bench_n2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], {k:v['value'] for k,v in (d.get('other_configs') or {}).items() if 'value' in v}, (d.get('other_configs') or {}).get('c1'))
"
