#!/bin/bash
tag=${1:-n4}; out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q -k "two_rank" > $out/${tag}_pytest_n2.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_n2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 20 --warmup 5 --no-other-scaling > $out/${tag}_n4.json 2> $out/${tag}_n4.err
MUSE_DEBUG_TIMING=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4 --steps 3 --warmup 5 --no-other-scaling > $out/${tag}_n4_dbg.json 2> $out/${tag}_n4_dbg.err
