#!/bin/bash
TAG=${1:-r03w}
mkdir -p gpurun_out
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --configs c3,c5 ) > gpurun_out/bench_1gpu_$TAG.log 2>&1; echo "bench 1 gpu rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
( time timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 5 --configs c3,c5,c4 ) > gpurun_out/bench_8gpu_$TAG.log 2>&1; echo "bench 8 gpus rc=$?"
