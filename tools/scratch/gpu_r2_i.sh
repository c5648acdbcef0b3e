#!/bin/bash
# round 2, session I (1 GPU): first run of the one-sweep GMRES kernels (fuse = sweep): granular parity tests, the
# fuse-parametrised solver tests, timing against block8, ncu launch list of k_sweep, kernel table (DG hoisted loads)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sweep.py -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2i_pytest_sweep.log 2>&1
echo "pytest sweep rc=$?" | tee -a gpurun_out/r2i_pytest_sweep.log; tail -5 gpurun_out/r2i_pytest_sweep.log
timeout 200 python tools/sweep_bench.py > gpurun_out/r2i_sweep_bench.txt 2>&1; echo "sweep_bench rc=$?"; cat gpurun_out/r2i_sweep_bench.txt | tail -8
timeout 420 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_kernels.py tests/test_gpu_orthogonality.py -m gpu -q --maxfail=40 --tb=short -k "sweep or orthogonal or dg or DG" > gpurun_out/r2i_pytest_fuse.log 2>&1
echo "pytest fuse rc=$?" | tee -a gpurun_out/r2i_pytest_fuse.log; tail -5 gpurun_out/r2i_pytest_fuse.log
timeout 240 ncu --kernel-name regex:k_sweep --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 64 --csv --log-file gpurun_out/r2i_sweep_launches.csv python tools/sweep_bench.py --short > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
timeout 200 python tools/quick_bench.py > gpurun_out/r2i_quick_bench.txt 2>&1; echo "quick_bench rc=$?"; grep -i "dg\|bratu1d\|block8" gpurun_out/r2i_quick_bench.txt
timeout 150 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x -k "one_strip or heat_periodic" > gpurun_out/r2i_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -15 gpurun_out/r2i_sanitizer.log
