#!/bin/bash
# quick gpurun: a pytest selection (-k expression in $1), output tail
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${2:-q}
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "$1" > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -40 $OUT/${TAG}_tests.log | cut -c1-400
