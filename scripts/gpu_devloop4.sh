#!/bin/bash
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "device_resident" > $out/${tag}_pytest_dev.log 2>&1
echo "pytest rc=$?" >> $out/${tag}_pytest_dev.log
export MUSE_FUSED_DRIVER=device
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $out/${tag}_c3_graph.json 2> $out/${tag}_c3_graph.err
timeout 200 python bench.py --d 512 --nsims 10000 --no-cpu-baseline > $out/${tag}_c2_graph.json 2> $out/${tag}_c2_graph.err
timeout 200 python bench.py --d 512 --nsims 100 --no-cpu-baseline > $out/${tag}_c1_graph.json 2> $out/${tag}_c1_graph.err
MUSE_D=512 MUSE_N=100 timeout 200 python scripts/host_overhead.py > $out/${tag}_host_c1_graph.log 2>&1
