#!/bin/bash
# round 2, session O (1 GPU): sweep kernel with a low-overhead row loop, one lane per column
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sweep.py -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2o_pytest_sweep.log 2>&1
echo "pytest sweep rc=$?" | tee -a gpurun_out/r2o_pytest_sweep.log; tail -5 gpurun_out/r2o_pytest_sweep.log
timeout 200 python tools/sweep_bench.py > gpurun_out/r2o_sweep_bench.txt 2>&1; echo "sweep_bench rc=$?"; tail -8 gpurun_out/r2o_sweep_bench.txt
AK_SWEEP_TEAM=0 timeout 200 python tools/sweep_bench.py --short > gpurun_out/r2o_sweep_bench_noteam.txt 2>&1; echo "noteam rc=$?"; tail -2 gpurun_out/r2o_sweep_bench_noteam.txt
timeout 240 ncu --kernel-name regex:k_sweep --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 21 --csv --log-file gpurun_out/r2o_sweep_launches.csv python tools/sweep_bench.py --short > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
AK_SWEEP_TEAM=0 timeout 240 ncu --kernel-name regex:k_sweep --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 20 --csv --log-file gpurun_out/r2o_sweep_launches_noteam.csv python tools/sweep_bench.py --short > gpurun_out/r2o_ncu2.log 2>&1; echo "ncu2 rc=$?"
