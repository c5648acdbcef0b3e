#!/bin/bash
# round 2, session N (1 GPU): sweep kernel with a low-overhead row loop, one lane per column
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sweep.py -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2n_pytest_sweep.log 2>&1
echo "pytest sweep rc=$?" | tee -a gpurun_out/r2n_pytest_sweep.log; tail -5 gpurun_out/r2n_pytest_sweep.log
timeout 200 python tools/sweep_bench.py > gpurun_out/r2n_sweep_bench.txt 2>&1; echo "sweep_bench rc=$?"; tail -8 gpurun_out/r2n_sweep_bench.txt
AK_SWEEP_NO_TEAM=1 timeout 200 python tools/sweep_bench.py --short > gpurun_out/r2n_sweep_bench_noteam.txt 2>&1; echo "noteam rc=$?"; tail -2 gpurun_out/r2n_sweep_bench_noteam.txt
timeout 240 ncu --kernel-name regex:k_sweep --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 21 --csv --log-file gpurun_out/r2n_sweep_launches.csv python tools/sweep_bench.py --short > gpurun_out/r2n_ncu.log 2>&1; echo "ncu rc=$?"
AK_SWEEP_SLACK=1 timeout 200 python tools/sweep_bench.py --short > gpurun_out/r2n_sweep_bench_slack1.txt 2>&1; tail -1 gpurun_out/r2n_sweep_bench_slack1.txt
AK_SWEEP_SLACK=3 timeout 200 python tools/sweep_bench.py --short > gpurun_out/r2n_sweep_bench_slack3.txt 2>&1; tail -1 gpurun_out/r2n_sweep_bench_slack3.txt
timeout 500 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_kernels.py tests/test_gpu_orthogonality.py tests/test_gpu_fullsize.py -m gpu -q --maxfail=20 --tb=short -k "sweep or orthogonal or c4 or C4 or cycle or c3 or C3" > gpurun_out/r2n_pytest_fuse.log 2>&1; tail -4 gpurun_out/r2n_pytest_fuse.log
