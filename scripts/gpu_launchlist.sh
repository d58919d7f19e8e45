#!/bin/bash
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
echo "rc=$?" >> $out/${tag}_bench_under_ncu.log
