#!/bin/bash
# weak-scaling sweep on one box: N = 1, 2, 4, 8 (needs gpurun --gpus 8).  Usage: bash scripts/gpu_scaling.sh <tag>
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $out/${tag}_n1.json 2> $out/${tag}_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 20 --warmup 3 > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err
done
