#!/bin/bash
# round 2, session M (1 GPU): sweep kernel with a low-overhead row loop, one lane per column
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sweep.py -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2m_pytest_sweep.log 2>&1
echo "pytest sweep rc=$?" | tee -a gpurun_out/r2m_pytest_sweep.log; tail -5 gpurun_out/r2m_pytest_sweep.log
timeout 200 python tools/sweep_bench.py > gpurun_out/r2m_sweep_bench.txt 2>&1; echo "sweep_bench rc=$?"; tail -8 gpurun_out/r2m_sweep_bench.txt
AK_SWEEP_NO_TEAM=1 timeout 200 python tools/sweep_bench.py --short > gpurun_out/r2m_sweep_bench_noteam.txt 2>&1; echo "noteam rc=$?"; tail -2 gpurun_out/r2m_sweep_bench_noteam.txt
timeout 240 ncu --kernel-name regex:k_sweep --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 21 --csv --log-file gpurun_out/r2m_sweep_launches.csv python tools/sweep_bench.py --short > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none --kernel-name regex:k_sweep --launch-skip 19 --launch-count 1 -o gpurun_out/r2m_sweep_k19 -f python tools/sweep_bench.py --short > gpurun_out/r2m_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r2m_sweep_k19.ncu-rep
