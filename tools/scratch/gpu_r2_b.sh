#!/bin/bash
# round 2, session B: GPU tests, kernel GB/s table, full default bench (with other_configs), ncu of the 1-D / DG kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=30 --deselect tests/test_gpu_multi.py > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
python tools/quick_bench.py > gpurun_out/r2b_quick_bench.txt 2>&1; echo "quick rc=$?"
cat gpurun_out/r2b_quick_bench.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench.err
python bench.py --config c3 --steps 5 --warmup 3 > gpurun_out/r2b_bench_c3.json 2>> gpurun_out/r2b_bench.err; echo "bench c3 rc=$?"
python tools/ncu_1d_driver.py 4 > gpurun_out/r2b_ncu1d_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_stencil1d|k_dg' -s 6 -c 6 -o gpurun_out/r02_ncu_1d python tools/ncu_1d_driver.py 4 > gpurun_out/r2b_ncu1d.log 2>&1
echo "ncu rc=$?"
