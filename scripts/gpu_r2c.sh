#!/bin/bash
# round 2: one-launch solve with lazy ẑ — tests, then A/B (lazy / stored ẑ / chain of launches) on C3, C2, C1, C4 and a stamp dump
tag=${1:-r2c}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
run() { # name, env..., -- args
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs $ARGS > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
}
ARGS="--steps 20 --warmup 5";                        run c3_lazy MUSE_PERSIST=1; run c3_stored MUSE_PERSIST=1 MUSE_LAZY=0; run c3_chain MUSE_PERSIST=0
ARGS="--d 512 --nsims 10000 --steps 50";             run c2_lazy MUSE_PERSIST=1; run c2_stored MUSE_PERSIST=1 MUSE_LAZY=0; run c2_chain MUSE_PERSIST=0
ARGS="--d 512 --nsims 100 --steps 100";              run c1_lazy MUSE_PERSIST=1; run c1_chain MUSE_PERSIST=0
ARGS="--family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3"; run c4_lazy MUSE_PERSIST=1; run c4_chain MUSE_PERSIST=0
for cfg in "65536 2048" "512 10000" "512 100"; do set -- $cfg
  MUSE_DEBUG_TIMING=1 MUSE_D=$1 MUSE_N=$2 timeout 120 python scripts/host_overhead.py > $out/${tag}_host_$1_$2.log 2>&1
done
