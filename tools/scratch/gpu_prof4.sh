#!/bin/bash
# round-1 evidence, final configuration (fuse = block4, un-normalised basis): launch list, --set full capture, both bench arms
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_b8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/launches_b8.csv $B > gpurun_out/ncu9.log 2>&1
$B > gpurun_out/plain_b8_2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mgs_block|k_stencil2d" -s 150 -c 14 -o gpurun_out/prof_b8 $B > gpurun_out/ncu10.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref5.json 2> gpurun_out/bench_ref5.err
python bench.py > gpurun_out/bench_b8_full.json 2> gpurun_out/bench_b8_full.err
python tools/quick_bench.py 8192 > gpurun_out/quick_bench5.log 2>&1
