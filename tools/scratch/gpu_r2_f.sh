#!/bin/bash
# round 2, session F (1 GPU): profiling evidence of the default path — ncu launch list, --set full capture of the dominant
# kernels, both bench arms, the other configs' own bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs"
$B > gpurun_out/r2f_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 520 -c 340 --csv --log-file gpurun_out/r02_launches_block8.csv $B > gpurun_out/r2f_ncu1.log 2>&1
echo "ncu launch list rc=$?"
$B > gpurun_out/r2f_plain2.log 2>&1 &&
ncu --set full --clock-control none -k regex:"k_mgs_block|k_stencil2d" -s 152 -c 8 -o gpurun_out/r02_prof_block8 $B > gpurun_out/r2f_ncu2.log 2>&1
echo "ncu set full rc=$?"
python tools/quick_bench.py > gpurun_out/r2f_quick_bench.txt 2>&1; tail -12 gpurun_out/r2f_quick_bench.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "bench rc=$?"
for c in c2 c5; do python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r2f_bench_$c.json 2> gpurun_out/r2f_bench_$c.err; echo "bench $c rc=$?"; done
python -c "
import json
for f in ('r2f_bench_n1','r2f_bench_reference','r2f_bench_c2','r2f_bench_c5'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('e2e',{}).get('value'), d.get('cpu_baseline',{}).get('value'))
"
