#!/bin/bash
# Evidence pass of the round's final build: full parity suite, smoke, bench lines, launch list and ncu captures.
TAG=${1:-r03z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.log 2>&1; echo "bench reference rc=$?"
timeout 300 python bench.py --workload c1 --configs none --no-cpu-baseline > gpurun_out/bench_c1_$TAG.log 2>&1; echo "bench c1 rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/ncu_chamfer.py 0 32 > gpurun_out/plain_ncu_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:chamfer_tc_kernel|chamfer_recover|chamfer_sort|chamfer_prune|chamfer_tc_plan' -s 12 -c 6 -f -o gpurun_out/prof_chamfer_$TAG \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full chamfer rc=$?"
tail -c 400 gpurun_out/pytest_$TAG.log
