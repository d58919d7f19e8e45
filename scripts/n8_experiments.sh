#!/bin/bash
out=gpurun_out; tag=$1
run() { name=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29900 + RANDOM % 50)) bench.py --gpus 8 --steps 20 --warmup 3 > $out/${tag}_$name.json 2> $out/${tag}_$name.err; }
run base A=1
run nosampler MUSE_BENCH_NO_SAMPLER=1
run ll NCCL_PROTO=LL MUSE_BENCH_NO_SAMPLER=1
run timing MUSE_DEBUG_TIMING=1 MUSE_BENCH_NO_SAMPLER=1
run base2 A=1
