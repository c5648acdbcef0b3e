#!/bin/bash
# 2 GPUs: new kernel tests, multi-GPU parity with the blocked sweep, 2-GPU bench (peer memory / NCCL)
set -x
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "collect or transpose or batched" 2>&1 | tail -25 > gpurun_out/pytest_f4.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -25 > gpurun_out/multi_b4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_b4_n2.json 2> gpurun_out/bench_b4_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-p2p > gpurun_out/bench_b4_n2_nccl.json 2> gpurun_out/bench_b4_n2_nccl.err
