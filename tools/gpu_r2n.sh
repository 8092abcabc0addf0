#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
python tools/ab_chamfer.py gpurun_tmp/libvpn_b200_r1.so volumetric-primitives-net_b200/lib/libvpn_b200.so > gpurun_out/ab_chamfer_$TAG.log 2>&1; cat gpurun_out/ab_chamfer_$TAG.log
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
tail -c 400 gpurun_out/pytest_$TAG.log
