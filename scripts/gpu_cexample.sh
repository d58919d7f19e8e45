#!/bin/bash
out=gpurun_out; mkdir -p $out
L=$PWD/museinference.jl_b200
gcc -std=c99 -O2 -Wall -I include examples/solve_funnel.c -o /tmp/solve_funnel -L $L -lmuse_b200 -lm -Wl,-rpath,$L > $out/r50_cexample.log 2>&1
timeout 30 /tmp/solve_funnel 4096 512 >> $out/r50_cexample.log 2>&1; echo "rc=$?" >> $out/r50_cexample.log
timeout 30 /tmp/solve_funnel 65536 2048 >> $out/r50_cexample.log 2>&1; echo "rc=$?" >> $out/r50_cexample.log
