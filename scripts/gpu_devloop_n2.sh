#!/bin/bash
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
for mode in device host; do
  MUSE_FUSED_DRIVER=$mode timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29700 + RANDOM % 50)) \
      bench.py --gpus 2 --steps 15 --warmup 5 > $out/${tag}_n2_$mode.json 2> $out/${tag}_n2_$mode.err
done
