#!/bin/bash
# full GPU test suite + smoke + C3/C1/C2 quick lines + host overhead with stamps
tag=${1:-f}; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 200 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs $ARGS > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err; }
ARGS="--steps 20 --warmup 5"; run c3 A=1
ARGS="--d 512 --nsims 10000 --steps 50";  run c2 A=1
ARGS="--d 512 --nsims 100 --steps 100";  run c1 A=1
for cfg in "65536 2048" "512 10000" "512 100"; do set -- $cfg
  MUSE_DEBUG_TIMING=1 MUSE_K=5 MUSE_D=$1 MUSE_N=$2 timeout 120 python scripts/host_overhead.py > $out/${tag}_host_$1_$2.log 2>&1
done
