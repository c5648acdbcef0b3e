#!/bin/bash
# un-normalised basis in the blocked sweeps: full GPU suite + bench
set -x
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu5.log
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench5_block4.json 2> gpurun_out/bench5_block4.err
python bench.py --steps 5 --fuse pair --no-cpu-baseline --no-e2e > gpurun_out/bench5_pair.json 2> gpurun_out/bench5_pair.err
