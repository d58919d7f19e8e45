#!/bin/bash
# N = 1, 4 and 8 on one 8-GPU box with the final sources (the full sweep is scripts/gpu_scale8.sh): strong scaling, weak as a sub-record
tag=${1:-s8f}; out=gpurun_out; mkdir -p $out
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $out/${tag}_n1.json 2> $out/${tag}_n1.err
for n in 8 4; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err
done
