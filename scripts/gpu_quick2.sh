#!/bin/bash
# quick check of a kernel change: the one-launch tests, then C3 / C4 / C2 bench lines
tag=${1:-q}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "one_launch or device_resident or full_muse or map_score_cold or tiny_and_ragged or user_start" > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs $ARGS > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err; }
ARGS="--steps 20 --warmup 5"; run c3 A=1; run c3_chain MUSE_PERSIST=0
ARGS="--family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3"; run c4 A=1
ARGS="--d 512 --nsims 10000 --steps 50";  run c2 A=1
