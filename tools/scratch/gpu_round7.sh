#!/bin/bash
# cached Gram entries in the blocked sweep: parity + bench
set -x
timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_fullsize.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gram.log
python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_gram.json 2> gpurun_out/bench_gram.err
python bench.py --steps 5 --fuse pair --no-cpu-baseline --no-e2e > gpurun_out/bench_gram_pair.json 2> gpurun_out/bench_gram_pair.err
