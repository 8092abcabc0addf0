#!/bin/bash
TAG=${1:-r05y}
mkdir -p gpurun_out
timeout 120 python tools/diag_chamfer.py 2>&1 | cut -c1-200 | grep -v "mismatches min1 0 idx1 0 min2 0 idx2 0" | tail -5
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --configs c3,c5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c2_$TAG.log; echo "bench rc=$?"
