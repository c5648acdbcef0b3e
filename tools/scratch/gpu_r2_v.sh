#!/bin/bash
# round 2, session V (2 GPUs): parity worker after making the sweep verdict collective
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 2 --master-port 29618 tests/multi_gpu_worker.py > gpurun_out/r02_multi_gpu_worker_2ranks.log 2>&1; echo "worker2 rc=$?"
grep -v "^$" gpurun_out/r02_multi_gpu_worker_2ranks.log | grep "multi-gpu\|MULTI\|Error\|error" | tail -16
