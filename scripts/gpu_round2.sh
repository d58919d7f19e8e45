#!/bin/bash
# Round-2 evidence session on one B200 (≈ 8 GPU-minutes): parity tests, smoke, the bench line (both arms, all configs), A/B of the
# round's design steps, ncu launch list, ncu full captures (one-launch solve with source, TMA GEMM), the plain-C example.
# Usage (repo root, under gpurun): bash scripts/gpu_round2.sh <tag>
tag=${1:-r02}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/${tag}_gpu.txt 2>&1; nproc >> $out/${tag}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 200 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
( time timeout 900 python bench.py ) > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
# A/B of the round's steps on C3 (and the chain of launches on C1, C2, C4)
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs $ARGS > $out/${tag}_ab_${name}.json 2> $out/${tag}_ab_${name}.err; }
ARGS="--steps 20 --warmup 5"
run c3_chain MUSE_PERSIST=0; run c3_stored_honest MUSE_LAZY=0 MUSE_LEAN=0; run c3_lean MUSE_LAZY=0; run c3_lazy MUSE_LEAN=0; run c3_lazy_lean_generic MUSE_FUNNEL_SPEC=0; run c3_default A=1
run c3_copy_node MUSE_HOSTWRITE=0; run c3_stream_sync MUSE_HOSTSPIN=0
ARGS="--d 512 --nsims 10000 --steps 50"; run c2_chain MUSE_PERSIST=0; run c2_copy_node MUSE_HOSTWRITE=0; run c2_stream_sync MUSE_HOSTSPIN=0; run c2_default A=1
ARGS="--d 512 --nsims 100 --steps 100"; run c1_chain MUSE_PERSIST=0; run c1_copy_node MUSE_HOSTWRITE=0; run c1_stream_sync MUSE_HOSTSPIN=0; run c1_default A=1
ARGS="--family twolayer --d 1024 --nsims 100 --steps 20"; run f4_default A=1
ARGS="--family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3"; run c4_chain MUSE_PERSIST=0; run c4_default A=1
ARGS="--family corrgauss --d 4096 --nsims 8192 --steps 3 --warmup 1"; run c5_cpasync MUSE_GEMM=cpasync; run c5_default A=1
timeout 300 python scripts/dgemm_bench.py > $out/${tag}_dgemm_tma_vs_cublas.txt 2>&1
MUSE_GEMM=cpasync timeout 300 python scripts/dgemm_bench.py > $out/${tag}_dgemm_cpasync_vs_cublas.txt 2>&1
timeout 300 python scripts/draws_ab.py > $out/${tag}_draws_ab.txt 2>&1
# ncu: launch list of the bench command, full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_persist -s 6 -c 1 -o $out/${tag}_persist_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_prof_persist.log 2>&1
python scripts/ncu_traffic.py $out/${tag}_persist_full.ncu-rep $out/${tag}_solver_traffic.json funnel 65536 2048 > $out/${tag}_traffic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma -s 2 -c 1 -o $out/${tag}_dgemm_tma_full \
    python scripts/dgemm_bench.py > $out/${tag}_prof_dgemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:philox_draws -s 3 -c 1 -o $out/${tag}_draws_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > $out/${tag}_prof_draws.log 2>&1
for cfg in "65536 2048" "512 10000" "512 100"; do set -- $cfg
  MUSE_DEBUG_TIMING=1 MUSE_K=5 MUSE_D=$1 MUSE_N=$2 timeout 120 python scripts/host_overhead.py > $out/${tag}_host_$1_$2.log 2>&1
done
L=$PWD/museinference.jl_b200
gcc -std=c99 -O2 -Wall -I include examples/solve_funnel.c -o /tmp/solve_funnel -L $L -lmuse_b200 -lm -Wl,-rpath,$L > $out/${tag}_cexample.log 2>&1
timeout 30 /tmp/solve_funnel 4096 512 >> $out/${tag}_cexample.log 2>&1; echo "rc=$?" >> $out/${tag}_cexample.log
timeout 30 /tmp/solve_funnel 65536 2048 >> $out/${tag}_cexample.log 2>&1; echo "rc=$?" >> $out/${tag}_cexample.log
