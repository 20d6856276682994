#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-x}
for mode in train; do
  echo "== bench_extra $mode"; timeout 900 python bench_extra.py $mode > $OUT/${TAG}_extra_$mode.json 2> $OUT/${TAG}_extra_$mode.err; echo "rc=$?"; cut -c1-600 $OUT/${TAG}_extra_$mode.json; tail -3 $OUT/${TAG}_extra_$mode.err
done
echo "== train tests"; timeout 1500 python -m pytest tests/test_gpu_train.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_train.log 2>&1; echo "rc=$?"; tail -15 $OUT/${TAG}_t_train.log | cut -c1-200
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
