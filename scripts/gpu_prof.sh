#!/bin/bash
# ncu evidence for the bench line: launch list of two steps + full capture of the four streaming launches of one step
# (cold pass, warm pass, fiducial, finite-difference pass).  Usage: bash scripts/gpu_prof.sh <tag>
tag=${1:-rX}
out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iso_stream -s 4 -c 4 -o $out/${tag}_stream_step_full \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $out/${tag}_prof_full.log 2>&1
