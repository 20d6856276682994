#!/bin/bash
# BASELINE configs 3-5 + the 128-chunk call shape on one GPU: one JSON line each.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-x}
for mode in train db sweep chunks; do
  echo "== bench_extra $mode"; timeout 900 python bench_extra.py $mode > $OUT/${TAG}_extra_$mode.json 2> $OUT/${TAG}_extra_$mode.err; echo "rc=$?"; cut -c1-700 $OUT/${TAG}_extra_$mode.json; tail -3 $OUT/${TAG}_extra_$mode.err
done
