#!/bin/bash
# deeper host speculation: full suite, mid-size timings, full-size bench
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest_gpu7.log
for n in 512 1024 2048; do echo "== N=$n"; python tools/quick_bench.py $n 2>&1 | grep -E "gmres"; done > gpurun_out/quick_sizes2.log 2>&1
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err
