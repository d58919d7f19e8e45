#!/bin/bash
# A/B of the completion-word wait of the one-launch solve (MUSE_HOSTSPIN=0: cudaStreamSynchronize) + the GPU test suite.
tag=${1:-spin}; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
run() { name=$1; shift; env "$@" timeout 300 python bench.py --no-cpu-baseline --no-extra-configs $ARGS > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err; }
ARGS="--steps 20 --warmup 5"; run c3_sync MUSE_HOSTSPIN=0; run c3_spin A=1
ARGS="--d 512 --nsims 10000 --steps 50"; run c2_sync MUSE_HOSTSPIN=0; run c2_spin A=1
ARGS="--d 512 --nsims 100 --steps 100"; run c1_sync MUSE_HOSTSPIN=0; run c1_spin A=1
for cfg in "65536 2048" "512 100"; do set -- $cfg
  MUSE_DEBUG_TIMING=1 MUSE_K=5 MUSE_D=$1 MUSE_N=$2 timeout 120 python scripts/host_overhead.py 2>&1 | grep "per solve\|whole call" | tail -3 > $out/${tag}_host_$1_$2.log
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for m in strong weak; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --scaling $m --no-cpu-baseline --no-extra-configs > $out/${tag}_n2_$m.json 2> $out/${tag}_n2_$m.err
  done
fi
