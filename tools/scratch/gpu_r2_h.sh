#!/bin/bash
# round 2, session H (2 GPUs): worker at 2 ranks (incl. the default-itmax case with uneven local sizes), 1-GPU suite,
# kernel table, C3 at the round-1 fusion level, 2-GPU bench (CPU leg with the other rank blocked in a socket wait)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29602 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_2ranks.log 2>&1; echo "worker2 rc=$?"
tail -3 gpurun_out/r02_multi_gpu_worker_2ranks.log
timeout 600 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py > gpurun_out/r2h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -3 gpurun_out/r2h_pytest.log
timeout 300 python tools/quick_bench.py > gpurun_out/r2h_quick_bench.txt 2>&1; tail -8 gpurun_out/r2h_quick_bench.txt
timeout 300 python bench.py --config c3 --fuse full --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_c3_full.json 2> gpurun_out/r2h_bench_c3_full.err; echo "c3 full rc=$?"
timeout 300 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2h_bench_c3_block8.json 2> gpurun_out/r2h_bench_c3_block8.err; echo "c3 b8 rc=$?"
timeout 400 $TR --nproc-per-node 2 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json
for f in ('r2h_bench_c3_full','r2h_bench_c3_block8','r2h_bench_n2'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('cpu_baseline'), (d.get('e2e') or {}).get('host_copy_GBs_all_ranks'))
"
