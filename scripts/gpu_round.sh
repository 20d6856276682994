#!/bin/bash
# One gpurun call: smoke, GPU parity tests (split so a faulting kernel cannot poison the rest),
# bench lines and ncu captures.  Everything lands in gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG=${1:-r1}
MODE=${2:-full}
echo "== smoke" ; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
tail -3 $OUT/${TAG}_smoke.log
echo "== kernels (non-TC)"; timeout 1200 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not tf32 and not auto" -p no:cacheprovider > $OUT/${TAG}_t_kernels.log 2>&1; echo "rc=$?"; tail -5 $OUT/${TAG}_t_kernels.log
echo "== kernels (TC)"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tf32 or auto" -p no:cacheprovider > $OUT/${TAG}_t_tc.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_t_tc.log
echo "== encoder (auto)"; timeout 1500 python -m pytest tests/test_gpu_encoder.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_t_enc_auto.log 2>&1; echo "rc=$?"; tail -8 $OUT/${TAG}_t_enc_auto.log
echo "== bench auto"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_auto.json 2> $OUT/${TAG}_bench_auto.err; echo "rc=$?"; cut -c1-400 $OUT/${TAG}_bench_auto.json; tail -3 $OUT/${TAG}_bench_auto.err
echo "== bench auto BN=256"; GRAFP_TC_BN=256 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_bn256.json 2> $OUT/${TAG}_bench_bn256.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_bn256.json
echo "== bench auto BN=64"; GRAFP_TC_BN=64 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_bn64.json 2> $OUT/${TAG}_bench_bn64.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_bn64.json
if [ "$MODE" = "full" ]; then
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_ref.json 2>&1; cut -c1-200 $OUT/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: aggregate + gemm_tc + knn"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mr_aggregate_staged -s 2 -c 1 -o $OUT/${TAG}_prof_agg -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_agg.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 4 -o $OUT/${TAG}_prof_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 2 -c 2 -o $OUT/${TAG}_prof_knn -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_knn.log 2>&1; echo "rc=$?"
fi
ls -la $OUT | grep ${TAG} | tail -30
