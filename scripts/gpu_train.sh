#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-t}
echo "== train"; timeout 1500 python -m pytest tests/test_gpu_train.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_train.log 2>&1; echo "rc=$?"; tail -40 $OUT/${TAG}_t_train.log
