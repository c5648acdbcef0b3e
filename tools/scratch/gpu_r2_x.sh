#!/bin/bash
# round 2, session X (1 GPU): the whole GPU suite on the final code
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=30 --tb=short --deselect tests/test_gpu_multi.py > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r2x_pytest.log; tail -8 gpurun_out/r2x_pytest.log
