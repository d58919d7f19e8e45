#!/bin/bash
# quick GPU check: parity tests + one bench line.  Usage: bash scripts/gpu_quick.sh <tag> [bench args]
tag=${1:-rX}; shift
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline "$@" > $out/${tag}_bench.json 2> $out/${tag}_bench.err
