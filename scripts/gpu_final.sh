#!/bin/bash
# Last GPU session of a round, ordered by priority (the call may be cut short): generator A/B, the new tests, the bench
# line, the full parity suite, smoke, then ncu evidence.  Usage (under gpurun): bash scripts/gpu_final.sh <tag>
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1; nproc >> $out/${tag}_gpu.txt
timeout 300 python scripts/draws_ab.py > $out/${tag}_draws_ab.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py -q -k "philox or fd_scores or transformed or profile_splits or fd_jacobian" > $out/${tag}_pytest_new.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_new.log
timeout 400 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:philox_draws -s 1 -c 1 -o $out/${tag}_draws_full \
    python scripts/draws_time.py > $out/${tag}_prof_draws.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
timeout 300 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_c2.json 2> $out/${tag}_c2.err
timeout 300 python bench.py --family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_c4.json 2> $out/${tag}_c4.err
