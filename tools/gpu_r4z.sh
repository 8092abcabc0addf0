#!/bin/bash
# Last check of the round's final tree: full parity suite, smoke, the default bench line, launch list.
TAG=${1:-r04z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
tail -c 300 gpurun_out/pytest_$TAG.log
