#!/bin/bash
# blocked Gram-Schmidt sweep (fuse = block4): parity tests + bench next to fuse = pair
set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu3.log
python bench.py --steps 5 --fuse pair --no-cpu-baseline --no-e2e > gpurun_out/bench3_pair.json 2> gpurun_out/bench3_pair.err
python bench.py --steps 5 --fuse block4 --no-cpu-baseline > gpurun_out/bench3_block4.json 2> gpurun_out/bench3_block4.err
python tools/quick_bench.py 8192 > gpurun_out/quick_bench3.log 2>&1
