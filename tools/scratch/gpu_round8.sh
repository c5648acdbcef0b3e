#!/bin/bash
# blocks of 8: parity + bench against blocks of 4
set -x
python bench.py --steps 5 --fuse block8 --no-cpu-baseline --no-e2e > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err
python bench.py --steps 5 --fuse block4 --no-cpu-baseline --no-e2e > gpurun_out/bench_b4b.json 2> gpurun_out/bench_b4b.err
