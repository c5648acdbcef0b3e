#!/bin/bash
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_pair.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 500 --csv --log-file gpurun_out/launches_pair.csv $B > gpurun_out/ncu3.log 2>&1
$B > gpurun_out/plain_pair2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mgs_pair|k_stencil2d" -s 200 -c 6 -o gpurun_out/prof_pair $B > gpurun_out/ncu4.log 2>&1
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref2.json 2> gpurun_out/bench_ref2.err
