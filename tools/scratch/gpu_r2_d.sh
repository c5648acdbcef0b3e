#!/bin/bash
# round 2, session D: GPU tests, kernel GB/s table, bench with and without the projection fusion
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
python tools/quick_bench.py > gpurun_out/r2d_quick_bench.txt 2>&1; echo "quick rc=$?"
tail -22 gpurun_out/r2d_quick_bench.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2d_bench.err
AK_NO_PROJ_FUSION=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-other-configs > gpurun_out/r2d_bench_nofusion.json 2>> gpurun_out/r2d_bench.err; echo "bench2 rc=$?"
python -c "
import json
for f in ('r2d_bench','r2d_bench_nofusion'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['value'], d['ms_per_step'], d['roofline']['frac_step'])
"
