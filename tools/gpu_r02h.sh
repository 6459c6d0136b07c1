#!/bin/bash
# 8-GPU scaling check: the driver's bench command at N=8 and the per-rank host cost (eager vs graph replay)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export LATTE_B200_NO_BUILD=1
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/h_bench_n$N.json 2> gpurun_out/h_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/gpu_hostcost.py > gpurun_out/h_hostcost_n$N.txt 2>&1
tail -n 1 gpurun_out/h_bench_n$N.json | cut -c1-2500; tail -n 5 gpurun_out/h_bench_n$N.err; grep rank gpurun_out/h_hostcost_n$N.txt | sort
