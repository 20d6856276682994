#!/bin/bash
# Final check of a build without the ncu passes: smoke, all GPU tests, both bench arms.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-fin}
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -1
echo "== tests"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_tests.log | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "rc=$?"; cut -c1-160 $OUT/${TAG}_bench_ref.json
echo "== bench (defaults)"; timeout 900 python bench.py > $OUT/${TAG}_bench_full.json 2> $OUT/${TAG}_bench_full.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_full.json; tail -2 $OUT/${TAG}_bench_full.err
