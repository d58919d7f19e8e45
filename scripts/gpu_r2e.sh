#!/bin/bash
# round 2: lean evaluation / lazy ẑ A/B inside the one-launch solve
tag=${1:-r2e}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs $ARGS > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err; }
ARGS="--steps 20 --warmup 5"
run c3_base MUSE_PERSIST=1; run c3_lean MUSE_LEAN=1; run c3_lazy MUSE_LAZY=1; run c3_lazylean MUSE_LAZY=1 MUSE_LEAN=1; run c3_chain MUSE_PERSIST=0
ARGS="--d 512 --nsims 10000 --steps 50";  run c2_base MUSE_PERSIST=1; run c2_lazylean MUSE_LAZY=1 MUSE_LEAN=1
ARGS="--d 512 --nsims 100 --steps 100";   run c1_base MUSE_PERSIST=1; run c1_lazylean MUSE_LAZY=1 MUSE_LEAN=1
ARGS="--family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3"; run c4_base MUSE_PERSIST=1; run c4_lazylean MUSE_LAZY=1 MUSE_LEAN=1
for cfg in "65536 2048" "512 10000" "512 100"; do set -- $cfg
  MUSE_DEBUG_TIMING=1 MUSE_K=5 MUSE_D=$1 MUSE_N=$2 timeout 120 python scripts/host_overhead.py > $out/${tag}_host_$1_$2.log 2>&1
done
