#!/bin/bash
# round 2, session W (1 GPU): evidence for the final state: whole GPU suite, smoke, both bench arms, ncu launch list of the
# bench command, ncu --set full of the sweep kernels, kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --tb=short --deselect tests/test_gpu_multi.py > gpurun_out/r2w_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2w_pytest.log; tail -4 gpurun_out/r2w_pytest.log
timeout 200 python __graft_entry__.py --smoke > gpurun_out/r2w_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2w_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.err; echo "reference rc=$?"; tail -c 600 gpurun_out/r2w_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r2w_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2w_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --kernel-name regex:k_sweep --launch-skip 7 --launch-count 13 -o gpurun_out/r2w_sweep_k8_20 -f python tools/sweep_bench.py --short > gpurun_out/r2w_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2w_sweep_k8_20.ncu-rep --page raw --csv > gpurun_out/r2w_sweep_k8_20_raw.csv 2>/dev/null; rm -f gpurun_out/r2w_sweep_k8_20.ncu-rep
timeout 200 python tools/quick_bench.py > gpurun_out/r2w_quick_bench.txt 2>&1; echo "quick_bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2w_bench_n1.json').read().strip().splitlines()[-1]); print('n1', d['value'], d['ms_per_step'], d.get('fuse'), (d.get('e2e') or {}).get('value'), d['roofline']['frac'], d['roofline'].get('frac_step'), d['gpu_launches'])
print({k:(v.get('value'), v.get('ms_per_step')) for k,v in (d.get('other_configs') or {}).items() if isinstance(v, dict)})
"
