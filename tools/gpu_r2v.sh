#!/bin/bash
TAG=${1:-r02v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q -x > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "pytest multigpu rc=$?" | tee -a gpurun_out/pytest_multigpu_$TAG.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline ) > gpurun_out/bench_1gpu_$TAG.log 2>&1; echo "bench 1 gpu rc=$?"
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  ( time timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_${N}gpu_$TAG.log 2>&1; echo "bench $N gpus rc=$?"
done
tail -3 gpurun_out/pytest_multigpu_$TAG.log
