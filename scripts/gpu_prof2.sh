#!/bin/bash
# ncu full capture (with source) of the one-launch solve at the C3 shape + per-line stall table
tag=${1:-p}; out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_persist -s 6 -c 1 -o $out/${tag}_persist_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_prof_full.log 2>&1
python scripts/ncu_traffic.py $out/${tag}_persist_full.ncu-rep $out/${tag}_solver_traffic.json funnel 65536 2048 > $out/${tag}_traffic.log 2>&1
ncu -i $out/${tag}_persist_full.ncu-rep --page source --csv > $out/${tag}_persist_source.csv 2> /dev/null
