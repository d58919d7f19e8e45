#!/bin/bash
# One GPU evidence session on the current code (≈ 6 GPU-minutes): parity tests, bench (both arms, all configs), switches A/B,
# ncu launch list, ncu full captures, timelines, the plain-C example.  Usage (repo root, under gpurun): bash scripts/gpu_round.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
timeout 120 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout 300 python bench.py --d 512 --nsims 100 --no-cpu-baseline > $out/${tag}_c1.json 2> $out/${tag}_c1.err
timeout 300 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_c2.json 2> $out/${tag}_c2.err
timeout 300 python bench.py --family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_c4.json 2> $out/${tag}_c4.err
timeout 300 python bench.py --family corrgauss --d 4096 --nsims 8192 --steps 3 --warmup 1 > $out/${tag}_c5.json 2> $out/${tag}_c5.err
# switches: outer loop (host arithmetic / device θ-step, eager / graph), normal generator (libm / table-driven)
for mode in host device; do
  MUSE_FUSED_DRIVER=$mode timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/${tag}_ab_c3_$mode.json 2> $out/${tag}_ab_c3_$mode.err
  MUSE_FUSED_DRIVER=$mode timeout 200 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_ab_c2_$mode.json 2> $out/${tag}_ab_c2_$mode.err
done
MUSE_OUTER_GRAPH=0 timeout 200 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_ab_c2_device_eager.json 2> $out/${tag}_ab_c2_device_eager.err
timeout 300 python scripts/draws_ab.py > $out/${tag}_draws_ab.txt 2>&1
# ncu: launch list of the bench command, full captures of the hot kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iso_stream -s 8 -c 4 -o $out/${tag}_stream_step_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_prof_stream.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:philox_draws -s 1 -c 1 -o $out/${tag}_draws_full \
    python scripts/draws_time.py > $out/${tag}_prof_draws.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"theta_step|cov_prep" -s 2 -c 2 -o $out/${tag}_outer_full \
    python bench.py --d 512 --nsims 10000 --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_prof_outer.log 2>&1
MUSE_N=2048 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dgemm|corr_iter" -s 2 -c 2 -o $out/${tag}_corr_full \
    python scripts/profile_corr.py > $out/${tag}_prof_corr.log 2>&1
timeout 300 python scripts/stream_timeline.py > $out/${tag}_stream_timeline.log 2>&1
MUSE_D=512 MUSE_N=10000 timeout 300 python scripts/host_overhead.py > $out/${tag}_host_c2.log 2>&1
timeout 300 python scripts/host_overhead.py > $out/${tag}_host_c3.log 2>&1
bash scripts/gpu_cexample.sh
