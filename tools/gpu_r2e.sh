#!/bin/bash
TAG=${1:-r02e}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_chamfer_prune.py tests/test_gpu_chamfer_fuzz.py -q -x > gpurun_out/pytest_prune_$TAG.log 2>&1; echo "pytest prune+fuzz rc=$?" | tee -a gpurun_out/pytest_prune_$TAG.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
timeout 300 python tools/ncu_chamfer.py 0 32 > gpurun_out/plain_ncu_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_chamfer_$TAG.csv \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_launches_chamfer_$TAG.log 2>&1
echo "ncu chamfer launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:chamfer_tc_kernel|chamfer_recover|chamfer_sort|chamfer_prune' -s 8 -c 6 -f -o gpurun_out/prof_chamfer_$TAG \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full chamfer rc=$?"
tail -c 400 gpurun_out/pytest_prune_$TAG.log; tail -c 700 gpurun_out/pytest_$TAG.log
