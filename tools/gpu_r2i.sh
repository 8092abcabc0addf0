#!/bin/bash
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chamfer_prune.py tests/test_gpu_chamfer_fuzz.py -q -x > gpurun_out/pytest_prune_$TAG.log 2>&1; echo "pytest prune+fuzz rc=$?"
timeout 300 python tools/bench_chamfer.py > gpurun_out/bench_chamfer_$TAG.log 2>&1; echo "bench chamfer rc=$?"
timeout 300 python tools/bench_chamfer.py c3 > gpurun_out/bench_chamfer_c3_$TAG.log 2>&1; echo "bench chamfer c3 rc=$?"
tail -3 gpurun_out/pytest_prune_$TAG.log
