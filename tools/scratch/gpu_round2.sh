#!/bin/bash
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python tools/quick_bench.py 8192 > gpurun_out/quick_bench.log 2>&1
for ry in 4 16 32; do echo "AK_RY=$ry"; AK_RY=$ry python tools/quick_bench.py 8192 2>&1 | grep -E "jvp bratu2d|residual bratu2d|fuse=full|fuse=pair"; done > gpurun_out/ry_sweep.log 2>&1
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err
