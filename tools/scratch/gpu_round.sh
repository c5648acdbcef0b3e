#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), ncu launch list + full capture of the dominant kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 700 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
$B > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_mgs_step -s 300 -c 3 -o gpurun_out/prof_mgs $B > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
