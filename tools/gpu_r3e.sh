#!/bin/bash
TAG=${1:-r03k}
mkdir -p gpurun_out
timeout 120 python tools/diag_chamfer.py 2>&1 | cut -c1-260 | grep -v "mismatches min1 0 idx1 0 min2 0 idx2 0" | tail -20
echo "diag rc=$?"
timeout 600 python -m pytest tests/test_gpu_chamfer_prune.py tests/test_gpu_chamfer_fuzz.py -q -x 2>&1 | tail -15
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "chamfer or end_to_end or step or pipeline" 2>&1 | tail -3
timeout 300 python bench.py --configs c3,c5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c2_$TAG.log; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu rc=$?"
