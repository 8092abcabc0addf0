#!/bin/bash
# Round-2 third pass (1 GPU): pruned tensor-core Chamfer filter - parity tests first (bounded by timeout), then bench.
TAG=${1:-r02c}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chamfer_prune.py -q -x > gpurun_out/pytest_prune_$TAG.log 2>&1; echo "pytest prune rc=$?" | tee -a gpurun_out/pytest_prune_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
timeout 300 python tools/bench_chamfer.py > gpurun_out/bench_chamfer_$TAG.log 2>&1; echo "bench chamfer rc=$?"
tail -c 1500 gpurun_out/pytest_prune_$TAG.log; tail -c 2500 gpurun_out/pytest_$TAG.log
