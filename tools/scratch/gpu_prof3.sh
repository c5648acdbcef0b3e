#!/bin/bash
# round-1 evidence for the blocked sweep (fuse = block4, default of bench.py): launch list, --set full capture, bench
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$B > gpurun_out/plain_b4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/launches_b4.csv $B > gpurun_out/ncu5.log 2>&1
$B > gpurun_out/plain_b4_2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_mgs_block|k_stencil2d" -s 150 -c 14 -o gpurun_out/prof_b4 $B > gpurun_out/ncu6.log 2>&1
python bench.py > gpurun_out/bench_b4_full.json 2> gpurun_out/bench_b4_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref3.json 2> gpurun_out/bench_ref3.err
