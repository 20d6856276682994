#!/bin/bash
# ncu evidence for the current build: launch list of one eager step (time + DRAM bytes per launch) and
# --set full captures of the aggregate, GEMM and kNN kernels (eager, real data, profiler range).
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-p}
GRAFP_NCU_RANGE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv python bench.py --no-graph --steps 1 --warmup 1 --batch 4096 --no-cpu-baseline --no-train --no-db --no-bf16 > $OUT/${TAG}_ncu_bench.log 2>&1; echo "launches rc=$?"
GRAFP_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mr_aggregate_staged -c 1 -o $OUT/${TAG}_prof_agg -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-db --no-bf16 > $OUT/${TAG}_ncu_agg.log 2>&1; echo "agg rc=$?"
GRAFP_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 24 -c 8 -o $OUT/${TAG}_prof_gemm -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-db --no-bf16 > $OUT/${TAG}_ncu_gemm.log 2>&1; echo "gemm rc=$?"
GRAFP_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:knn_tc_kernel -c 5 -o $OUT/${TAG}_prof_knn -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-db --no-bf16 > $OUT/${TAG}_ncu_knn.log 2>&1; echo "knn rc=$?"
GRAFP_NCU_RANGE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:ffn_fused_kernel -c 4 -o $OUT/${TAG}_prof_ffn -f python bench.py --no-graph --steps 1 --warmup 1 --no-cpu-baseline --no-train --no-db --no-bf16 > $OUT/${TAG}_ncu_ffn.log 2>&1; echo "ffn rc=$?"
