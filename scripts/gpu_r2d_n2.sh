#!/bin/bash
# round 2, two GPUs: the 2-rank bit-identity test (peer-mapped exchange and NCCL), then strong-scaling bench lines with both exchanges
tag=${1:-r2d}; out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "two_rank or nccl_shardpool" > $out/${tag}_pytest_n2.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_n2.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
MUSE_DEBUG_TIMING= timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-other-scaling > $out/${tag}_n2_strong_p2p.json 2> $out/${tag}_n2_strong_p2p.err
MUSE_EXCHANGE=nccl timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-other-scaling > $out/${tag}_n2_strong_nccl.json 2> $out/${tag}_n2_strong_nccl.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 5 --scaling weak --no-other-scaling > $out/${tag}_n2_weak_p2p.json 2> $out/${tag}_n2_weak_p2p.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $out/${tag}_n1.json 2> $out/${tag}_n1.err
