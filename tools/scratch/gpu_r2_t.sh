#!/bin/bash
# round 2, session T (1 GPU): 1-D sweep kernels (Bratu 1-D, heat 1-D, DG): parity, then the C2 / C5 bench lines
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sweep.py -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2t_pytest_sweep.log 2>&1
echo "pytest sweep rc=$?" | tee -a gpurun_out/r2t_pytest_sweep.log; tail -25 gpurun_out/r2t_pytest_sweep.log
timeout 420 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_kernels.py tests/test_gpu_orthogonality.py -m gpu -q --maxfail=40 --tb=short -k "sweep or orthogonal" > gpurun_out/r2t_pytest_fuse.log 2>&1
echo "pytest fuse rc=$?" | tee -a gpurun_out/r2t_pytest_fuse.log; tail -12 gpurun_out/r2t_pytest_fuse.log
for cfg in c2 c5; do
timeout 300 python bench.py --config $cfg --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2t_bench_$cfg.json 2> gpurun_out/r2t_bench_$cfg.err; echo "bench $cfg rc=$?"
timeout 300 python bench.py --config $cfg --fuse block8 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2t_bench_${cfg}_block8.json 2> gpurun_out/r2t_bench_${cfg}_block8.err
done
python -c "
import json
for f in ('c2','c2_block8','c5','c5_block8'):
    d=json.loads(open('gpurun_out/r2t_bench_%s.json'%f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('frac_step'))
"
