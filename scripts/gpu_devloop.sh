#!/bin/bash
# device-resident outer loop: parity tests, A/B bench against the host loop, then the whole suite.  Usage: bash scripts/gpu_devloop.sh <tag>
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "device_resident" > $out/${tag}_pytest_dev.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_dev.log
for mode in host device; do
  MUSE_FUSED_DRIVER=$mode timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/${tag}_c3_$mode.json 2> $out/${tag}_c3_$mode.err
  MUSE_FUSED_DRIVER=$mode timeout 200 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_c2_$mode.json 2> $out/${tag}_c2_$mode.err
  MUSE_FUSED_DRIVER=$mode timeout 200 python bench.py --d 512 --nsims 100 --no-cpu-baseline > $out/${tag}_c1_$mode.json 2> $out/${tag}_c1_$mode.err
done
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest.log
