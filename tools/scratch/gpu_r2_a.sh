#!/bin/bash
# round 2, session A: full GPU test suite + a short bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -x --deselect tests/test_gpu_multi.py > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/r2a_bench.json
