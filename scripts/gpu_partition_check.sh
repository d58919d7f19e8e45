#!/bin/bash
# Partition invariance end to end: the same total problem on 1 GPU and sharded over 2 GPUs must give the same θ̂, σ.
out=gpurun_out; tag=$1
for fam in "funnel 8192" "hiergauss 6000" "corrgauss 512"; do
  set -- $fam
  [ -s $out/${tag}_$1_n1.json ] || python bench.py --gpus 1 --family $1 --d $2 --nsims 400 --scaling strong --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_$1_n1.json 2> $out/${tag}_$1_n1.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 \
      bench.py --gpus 2 --family $1 --dim $2 --nsims 400 --scaling strong --steps 3 --warmup 3 > $out/${tag}_$1_n2.json 2> $out/${tag}_$1_n2.err
done
