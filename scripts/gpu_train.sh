#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-t}
timeout 900 python scripts/diag_train.py > $OUT/${TAG}_diag_train.log 2>&1; tail -5 $OUT/${TAG}_diag_train.log
