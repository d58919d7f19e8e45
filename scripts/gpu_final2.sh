#!/bin/bash
# Remaining evidence of the final code: C1, C5, the reference arm, host overhead.  Usage: bash scripts/gpu_final2.sh <tag>
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
timeout 300 python bench.py --d 512 --nsims 100 --no-cpu-baseline > $out/${tag}_c1.json 2> $out/${tag}_c1.err
timeout 400 python bench.py --family corrgauss --d 4096 --nsims 8192 --steps 3 --warmup 1 > $out/${tag}_c5.json 2> $out/${tag}_c5.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
MUSE_D=512 MUSE_N=10000 timeout 200 python scripts/host_overhead.py > $out/${tag}_host_c2.log 2>&1
timeout 200 python scripts/host_overhead.py > $out/${tag}_host_c3.log 2>&1
