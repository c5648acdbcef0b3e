#!/bin/bash
# 8-GPU box: multi-GPU parity at 4 and 8 ranks + weak scaling of the final configuration (run with gpurun --gpus 8)
set -x
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.log
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "4 or 8" 2>&1 | tail -30 > gpurun_out/multi8_final.log
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/scale2_n1.json 2> gpurun_out/scale2_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29620+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale2_n$n.json 2> gpurun_out/scale2_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 8 --steps 5 --warmup 3 --no-p2p --no-e2e > gpurun_out/scale2_n8_nccl.json 2> gpurun_out/scale2_n8_nccl.err
