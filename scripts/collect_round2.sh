#!/bin/bash
# Copy the evidence of one scripts/gpu_round2.sh session from gpurun_out/ (scratch) into profiles/ (tracked); run at the repo root.
# Usage: bash scripts/collect_round2.sh [tag]      (needs ncu and cuobjdump, no GPU)
tag=${1:-r02}; o=gpurun_out; p=profiles
for f in $o/${tag}_ab_*.json; do cp $f $p/$(basename $f); done
cp $o/${tag}_bench.json $p/${tag}_bench.json
cp $o/${tag}_bench_ref.json $p/${tag}_bench_reference_arm.json
cp $o/${tag}_pytest.log $p/${tag}_gpu_tests.txt
cp $o/${tag}_cexample.log $p/${tag}_c_example.txt
cp $o/${tag}_launches.csv $p/${tag}_launches_bench_steps2.csv
cp $o/${tag}_dgemm_tma_vs_cublas.txt $o/${tag}_dgemm_cpasync_vs_cublas.txt $p/
for c in 65536_2048 512_10000 512_100; do cp $o/${tag}_host_$c.log $p/${tag}_host_overhead_$c.txt; done
cp $o/${tag}_solver_traffic.json $p/${tag}_solver_traffic.json
{ head -6 $p/${tag}_draws_ab.txt; grep "^impl" $o/${tag}_draws_ab.txt; } > /tmp/_draws_ab.txt && cp /tmp/_draws_ab.txt $p/${tag}_draws_ab.txt
python scripts/ncu_summary.py $o/${tag}_draws_full.ncu-rep $p/${tag}_draws_ncu_full
python scripts/ncu_summary.py $o/${tag}_persist_full.ncu-rep $p/${tag}_persist_ncu_full
python scripts/ncu_summary.py $o/${tag}_dgemm_tma_full.ncu-rep $p/${tag}_dgemm_tma_ncu_full
python scripts/sass_census.py $p/${tag}_sass_census.txt
