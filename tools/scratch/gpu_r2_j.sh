#!/bin/bash
# round 2, session J (1 GPU): sweep kernel with independent warps + team decomposition: parity, timing, in-kernel cycle
# counters, ncu launch list
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sweep.py -m gpu -q --maxfail=40 --tb=short > gpurun_out/r2j_pytest_sweep.log 2>&1
echo "pytest sweep rc=$?" | tee -a gpurun_out/r2j_pytest_sweep.log; tail -5 gpurun_out/r2j_pytest_sweep.log
timeout 200 python tools/sweep_bench.py > gpurun_out/r2j_sweep_bench.txt 2>&1; echo "sweep_bench rc=$?"; tail -8 gpurun_out/r2j_sweep_bench.txt
AK_SWEEP_NO_TEAM=1 timeout 200 python tools/sweep_bench.py > gpurun_out/r2j_sweep_bench_noteam.txt 2>&1; echo "noteam rc=$?"; tail -8 gpurun_out/r2j_sweep_bench_noteam.txt
AK_SWEEP_DEBUG=1 timeout 200 python tools/sweep_bench.py --short > gpurun_out/r2j_sweep_debug.txt 2>&1; echo "debug rc=$?"; grep k_sweep gpurun_out/r2j_sweep_debug.txt | head -24
timeout 240 ncu --kernel-name regex:k_sweep --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 21 --csv --log-file gpurun_out/r2j_sweep_launches.csv python tools/sweep_bench.py --short > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"
for mb in 0 6 8; do AK_DG_MB=$mb timeout 100 python tools/dg_bench.py >> gpurun_out/r2j_dg_bench.txt 2>&1; done; cat gpurun_out/r2j_dg_bench.txt
for sk in 1 3 4 8 15; do echo "AK_SWEEP_SKIP=$sk"; AK_SWEEP_SKIP=$sk timeout 100 python tools/sweep_bench.py --short 2>&1 | tail -1; done > gpurun_out/r2j_sweep_skip.txt 2>&1; cat gpurun_out/r2j_sweep_skip.txt
