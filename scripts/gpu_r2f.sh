#!/bin/bash
# round 2: ncu evidence of the one-launch solve (launch list + full capture with source), then the default bench line
tag=${1:-r2f}; out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_persist -s 6 -c 1 -o $out/${tag}_persist_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_prof_full.log 2>&1
python scripts/ncu_traffic.py $out/${tag}_persist_full.ncu-rep $out/${tag}_solver_traffic.json funnel 65536 2048 > $out/${tag}_traffic.log 2>&1
ncu -i $out/${tag}_persist_full.ncu-rep --page source --csv > $out/${tag}_persist_source.csv 2> /dev/null
( time timeout 900 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
