#!/bin/bash
# TMA-fed DGEMM: parity tests, timing against cuBLAS and against the cp.async form, then the corrgauss tests and the C5 bench line
tag=${1:-g}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "dgemm or corrgauss or implicit" > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
timeout 300 python scripts/dgemm_bench.py > $out/${tag}_dgemm_tma.txt 2>&1
MUSE_GEMM=cpasync timeout 300 python scripts/dgemm_bench.py > $out/${tag}_dgemm_cpasync.txt 2>&1
timeout 600 python bench.py --family corrgauss --d 4096 --nsims 8192 --steps 3 --warmup 1 --no-cpu-baseline > $out/${tag}_c5.json 2> $out/${tag}_c5.err
MUSE_GEMM=cpasync timeout 600 python bench.py --family corrgauss --d 4096 --nsims 8192 --steps 3 --warmup 1 --no-cpu-baseline > $out/${tag}_c5_cpasync.json 2> $out/${tag}_c5_cpasync.err
