#!/bin/bash
# 1/2/4/8-GPU weak-scaling session on one box (run with gpurun --gpus 8)
set -x
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus.log
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -30 > gpurun_out/multi8.log
python bench.py --steps 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
