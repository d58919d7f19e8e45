#!/bin/bash
# round 2, first session: new parity tests at size, the reworked bench line, launch list + full capture for the traffic record
tag=${1:-r2a}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iso_stream -s 12 -c 4 -o $out/${tag}_stream_step_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_prof_full.log 2>&1
python scripts/ncu_traffic.py $out/${tag}_stream_step_full.ncu-rep $out/${tag}_solver_traffic.json funnel 65536 2048 > $out/${tag}_traffic.log 2>&1
