#!/bin/bash
# gpurun --gpus N: NCCL train-step check and the bench line at N ranks
set -u
N=${1:-4}; TAG=${2:-mn}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
echo "== multi-gpu train check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/multi_gpu_check.py > $OUT/${TAG}_multi_check.log 2>&1; echo "rc=$?"; grep MULTI_GPU_CHECK $OUT/${TAG}_multi_check.log | cut -c1-400 || tail -20 $OUT/${TAG}_multi_check.log
echo "== bench N=$N"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "rc=$?"; cut -c1-260 $OUT/${TAG}_bench_n$N.json; tail -2 $OUT/${TAG}_bench_n$N.err
