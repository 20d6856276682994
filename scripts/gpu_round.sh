#!/bin/bash
# Full gpurun: smoke, all GPU tests, bench lines (both arms), ncu launch list + full captures.
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r1}
echo "== smoke" ; timeout 600 python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
echo "== tests"; timeout 2400 python -m pytest tests -q -m gpu -p no:cacheprovider > $OUT/${TAG}_tests.log 2>&1; echo "rc=$?"; tail -12 $OUT/${TAG}_tests.log | cut -c1-200
echo "== bench (default)"; timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
echo "== bench (eager, no graph)"; timeout 900 python bench.py --no-graph --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_eager.json 2> $OUT/${TAG}_bench_eager.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_eager.json; tail -2 $OUT/${TAG}_bench_eager.err
echo "== bench (3xtf32)"; timeout 900 python bench.py --engine 3xtf32 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_3xtf32.json 2> $OUT/${TAG}_bench_3xtf32.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_3xtf32.json
echo "== bench (bf16)"; timeout 900 python bench.py --engine bf16 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_bf16.json 2> $OUT/${TAG}_bench_bf16.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_bf16.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_ref.json 2>&1; cut -c1-200 $OUT/${TAG}_bench_ref.json
echo "== ncu launch list of one eager step (profiler range), with DRAM bytes per launch"
GRAFP_NCU_RANGE=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python bench.py --no-graph --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "rc=$?"
echo "== ncu full: aggregate + gemm_tc + knn (eager, real data)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mr_aggregate_staged -s 12 -c 1 -o $OUT/${TAG}_prof_agg -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_agg.log 2>&1; echo "rc=$?"
GRAFP_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 28 -c 8 -o $OUT/${TAG}_prof_gemm -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "rc=$?"
GRAFP_NCU_RANGE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 5 -o $OUT/${TAG}_prof_knn -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_knn.log 2>&1; echo "rc=$?"
