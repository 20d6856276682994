#!/bin/bash
# gpurun --gpus 2: NCCL train-step check + 1/2-GPU bench lines (both arms)
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-m}
nvidia-smi -L | head -4
echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > $OUT/${TAG}_multi_tests.log 2>&1; echo "rc=$?"; tail -3 $OUT/${TAG}_multi_tests.log
echo "== multi-gpu train check"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/multi_gpu_check.py > $OUT/${TAG}_multi_check.log 2>&1; echo "rc=$?"; grep MULTI_GPU_CHECK $OUT/${TAG}_multi_check.log || tail -20 $OUT/${TAG}_multi_check.log
echo "== bench N=1"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_n1.json
echo "== bench N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err; echo "rc=$?"; cut -c1-300 $OUT/${TAG}_bench_n2.json; tail -3 $OUT/${TAG}_bench_n2.err
echo "== bench reference N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $OUT/${TAG}_bench_ref_n2.json 2> $OUT/${TAG}_bench_ref_n2.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_ref_n2.json
