#!/bin/bash
out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize.py > $out/r02_sanitize_$tool.log 2>&1
  echo "rc=$?" >> $out/r02_sanitize_$tool.log
done
