#!/bin/bash
# round 2, session Q (1 GPU): whole GPU suite with fuse = sweep as the library default, smoke, default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --tb=short --deselect tests/test_gpu_multi.py > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2q_pytest.log; tail -6 gpurun_out/r2q_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_n1.json 2> gpurun_out/r2q_bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/r2q_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2q_bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d.get('fuse'), (d.get('e2e') or {}).get('value'), d['roofline']['frac'], d['roofline'].get('frac_step'), d['gpu_launches']); print({k:(v.get('value'), v.get('ms_per_step')) for k,v in (d.get('other_configs') or {}).items()} if isinstance(d.get('other_configs'), dict) else d.get('other_configs'))
"
