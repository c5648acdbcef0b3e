#!/bin/bash
# 2 GPUs, blocks of 8 as the default: multi-GPU parity (NCCL and peer memory) and the 2-GPU bench
set -x
timeout 800 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -25 > gpurun_out/multi_b8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29833 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_b8_n2.json 2> gpurun_out/bench_b8_n2.err
