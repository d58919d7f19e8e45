#!/bin/bash
# round 2, one-launch solve: its tests first (bounded), then the A/B against the chain of launches on C3 / C2 / C1
tag=${1:-r2b}; out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "one_launch or device_resident or full_muse or profile_splits" > $out/${tag}_pytest_persist.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_persist.log
for p in 1 0; do
  MUSE_PERSIST=$p timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > $out/${tag}_c3_persist$p.json 2> $out/${tag}_c3_persist$p.err
  MUSE_PERSIST=$p timeout 300 python bench.py --d 512 --nsims 10000 --steps 50 --no-cpu-baseline > $out/${tag}_c2_persist$p.json 2> $out/${tag}_c2_persist$p.err
  MUSE_PERSIST=$p timeout 300 python bench.py --d 512 --nsims 100 --steps 100 --no-cpu-baseline > $out/${tag}_c1_persist$p.json 2> $out/${tag}_c1_persist$p.err
done
MUSE_PERSIST=1 timeout 300 python bench.py --family hiergauss --d 100000 --nsims 4096 --steps 10 --warmup 3 --no-cpu-baseline > $out/${tag}_c4_persist1.json 2> $out/${tag}_c4_persist1.err
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
