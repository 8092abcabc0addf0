#!/bin/bash
# Round-2 second pass (2 GPUs): all parity tests incl. multi-GPU / fuzz / train-loop integration, the new bench line with
# its configs block on 1 GPU, and the 2-GPU scaling line with the self-synchronising NVLS all-reduce (variants).
TAG=${1:-r02b}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
( time python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.log 2>&1; echo "bench reference rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time $TR bench.py --gpus 2 ) > gpurun_out/bench_2gpu_auto_$TAG.log 2>&1; echo "2gpu auto rc=$?"
VPN_ALLREDUCE=nvls $TR bench.py --gpus 2 --configs none > gpurun_out/bench_2gpu_nvls_$TAG.log 2>&1; echo "2gpu nvls rc=$?"
VPN_ALLREDUCE=nvls VPN_BENCH_AR_IN_GRAPH=0 $TR bench.py --gpus 2 --configs none > gpurun_out/bench_2gpu_nvls_nograph_$TAG.log 2>&1; echo "2gpu nvls (not in graph) rc=$?"
VPN_ALLREDUCE=nccl $TR bench.py --gpus 2 --configs none > gpurun_out/bench_2gpu_nccl_$TAG.log 2>&1; echo "2gpu nccl rc=$?"
VPN_ALLREDUCE=nvls VPN_BENCH_NO_CLOCKS=1 $TR bench.py --gpus 2 --configs none > gpurun_out/bench_2gpu_nvls_noclk_$TAG.log 2>&1; echo "2gpu nvls noclk rc=$?"
tail -c 2500 gpurun_out/pytest_$TAG.log
