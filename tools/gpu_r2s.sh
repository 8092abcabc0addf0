#!/bin/bash
# 8-GPU pass: multi-GPU parity test on all ranks, then the scaling lines the driver's SCALE run takes (N = 8, 4, 2).
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q -x > gpurun_out/pytest_multigpu_$TAG.log 2>&1; echo "pytest multigpu rc=$?" | tee -a gpurun_out/pytest_multigpu_$TAG.log
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  ( time timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_${N}gpu_$TAG.log 2>&1; echo "bench $N gpus rc=$?"
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
VPN_ALLREDUCE=nccl timeout 300 $TR bench.py --gpus 8 --configs none > gpurun_out/bench_8gpu_nccl_$TAG.log 2>&1; echo "bench 8 gpus nccl rc=$?"
VPN_BENCH_AR_IN_GRAPH=0 timeout 300 $TR bench.py --gpus 8 --configs none > gpurun_out/bench_8gpu_nograph_$TAG.log 2>&1; echo "bench 8 gpus ar-not-in-graph rc=$?"
( time timeout 300 $TR bench.py --gpus 8 --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_8gpu_ref_$TAG.log 2>&1; echo "bench 8 gpus reference arm rc=$?"
tail -3 gpurun_out/pytest_multigpu_$TAG.log
