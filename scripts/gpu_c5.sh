#!/bin/bash
tag=${1:-c5}; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q -k "dgemm or corrgauss or implicit or two_rank" > $out/${tag}_pytest.log 2>&1; echo "rc=$?" >> $out/${tag}_pytest.log
timeout 600 python bench.py --family corrgauss --d 4096 --nsims 8192 --steps 3 --warmup 1 --no-cpu-baseline > $out/${tag}_c5.json 2> $out/${tag}_c5.err
