#!/bin/bash
# 1/2/4/8-GPU sweep on one box: strong scaling (default, the north-star split) with the weak figure as a sub-record; N=2 parity test first
tag=${1:-s8}; out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "two_rank" > $out/${tag}_pytest_n2.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_n2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $out/${tag}_n1.json 2> $out/${tag}_n1.err
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 > $out/${tag}_n$n.json 2> $out/${tag}_n$n.err
done
MUSE_EXCHANGE=nccl timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 5 --no-other-scaling > $out/${tag}_n8_nccl.json 2> $out/${tag}_n8_nccl.err
MUSE_DEBUG_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 5 --no-other-scaling > $out/${tag}_n8_dbg.json 2> $out/${tag}_n8_dbg.err
