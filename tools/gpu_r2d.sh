#!/bin/bash
# Round-2 pass: pruned Chamfer (builder skip, hierarchical bounds, faster sort) + rasteriser group-box cull.
TAG=${1:-r02d}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chamfer_prune.py -q -x > gpurun_out/pytest_prune_$TAG.log 2>&1; echo "pytest prune rc=$?" | tee -a gpurun_out/pytest_prune_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
timeout 300 python tools/bench_chamfer.py > gpurun_out/bench_chamfer_$TAG.log 2>&1; echo "bench chamfer rc=$?"
timeout 300 python tools/ncu_chamfer.py 0 32 > gpurun_out/plain_ncu_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_chamfer_$TAG.csv \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_launches_chamfer_$TAG.log 2>&1
echo "ncu chamfer launches rc=$?"
timeout 300 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/plain_c3_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c3_$TAG.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_launches_c3_$TAG.log 2>&1
echo "ncu c3 launches rc=$?"
tail -c 600 gpurun_out/pytest_prune_$TAG.log; tail -c 1500 gpurun_out/pytest_$TAG.log
