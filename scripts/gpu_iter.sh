#!/bin/bash
# Iteration gpurun: all GPU tests (fail fast), determinism diagnostic, one bench line with per-shape kernel times.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-it}
echo "== tests"; timeout 1200 python -m pytest tests -q -m gpu -x -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -6 $OUT/${TAG}_tests.log | cut -c1-300
echo "== determinism"; timeout 300 python scripts/diag_batch.py 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; cut -c1-180 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
