#!/bin/bash
tag=${1:-rX}; out=gpurun_out; mkdir -p $out
export MUSE_FUSED_DRIVER=device
MUSE_DEBUG_TIMING=1 MUSE_D=512 MUSE_N=100 timeout 200 python scripts/host_overhead.py > $out/${tag}_dbg_c1.log 2>&1
