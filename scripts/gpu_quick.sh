#!/bin/bash
# Short gpurun: TC kernel tests, encoder tests, bench, ncu of the two tensor-core kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-q}
echo "== kernels"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_kernels.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_t_kernels.log
echo "== encoder"; timeout 1500 python -m pytest tests/test_gpu_encoder.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_enc.log 2>&1; echo "rc=$?"; tail -6 $OUT/${TAG}_t_enc.log
echo "== bench auto"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_auto.json 2> $OUT/${TAG}_bench_auto.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_auto.json; tail -3 $OUT/${TAG}_bench_auto.err
echo "== bench BN=256"; GRAFP_TC_BN=256 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_bn256.json 2> $OUT/${TAG}_bench_bn256.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_bn256.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 4 -o $OUT/${TAG}_prof_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 2 -o $OUT/${TAG}_prof_knn -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_knn.log 2>&1; echo "rc=$?"
echo "== train"; timeout 1500 python -m pytest tests/test_gpu_train.py -q -m gpu -p no:cacheprovider -x > $OUT/${TAG}_t_train.log 2>&1; echo "rc=$?"; tail -25 $OUT/${TAG}_t_train.log
