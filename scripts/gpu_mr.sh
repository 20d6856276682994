#!/bin/bash
# Short gpurun for the fused MRConv -> fc2 kernel: its bit-exact test, the fused-FFN test, encoder tests, A/B bench.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-mr}
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "fused" > $OUT/${TAG}_t_kernels.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_t_kernels.log | cut -c1-300
echo "== encoder"; timeout 1500 python -m pytest tests/test_gpu_encoder.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_enc.log 2>&1; echo "rc=$?"; tail -6 $OUT/${TAG}_t_enc.log | cut -c1-300
echo "== bench fused"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-bf16 --no-db > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench two-GEMM"; GRAFP_NO_MR_FUSED=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-train --no-bf16 --no-db > $OUT/${TAG}_bench_off.json 2> $OUT/${TAG}_bench_off.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_off.json
python scripts/cmp_bench.py $OUT/${TAG}_bench_off.json $OUT/${TAG}_bench.json | head -45
