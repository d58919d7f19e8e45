#!/bin/bash
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
timeout 120 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1
timeout 100 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_c2.json 2> $out/${tag}_c2.err
timeout 100 python bench.py --d 512 --nsims 100 --no-cpu-baseline > $out/${tag}_c1.json 2> $out/${tag}_c1.err
timeout 100 python bench.py --family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_c4.json 2> $out/${tag}_c4.err
