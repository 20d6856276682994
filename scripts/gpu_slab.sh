#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-s}
for mb in 0 16 32 48 64 96; do
  echo "== slab $mb MB"; GRAFP_FFN_SLAB_MB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_slab$mb.json 2> $OUT/${TAG}_bench_slab$mb.err; echo "rc=$?"; cut -c1-140 $OUT/${TAG}_bench_slab$mb.json; tail -2 $OUT/${TAG}_bench_slab$mb.err
done
echo "== encoder tests with slabs"; timeout 900 python -m pytest tests/test_gpu_encoder.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_enc.log 2>&1; echo "rc=$?"; tail -4 $OUT/${TAG}_t_enc.log
