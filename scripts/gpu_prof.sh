#!/bin/bash
# ncu launch list + full capture of the solver kernels.  Usage: bash scripts/gpu_prof.sh <tag>
tag=${1:-rX}
out=gpurun_out; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iso_stream -s 0 -c 3 -o $out/${tag}_stream_full \
    python scripts/profile_solver.py > $out/${tag}_prof_full.log 2>&1
