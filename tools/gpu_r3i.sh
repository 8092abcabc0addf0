#!/bin/bash
TAG=${1:-r03s}
mkdir -p gpurun_out
timeout 300 python tools/ncu_chamfer.py 0 32 > gpurun_out/plain_ncu_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:chamfer_tc_kernel|chamfer_recover|chamfer_sort|chamfer_prune|chamfer_tc_plan' -s 12 -c 6 -f -o gpurun_out/prof_chamfer_$TAG \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full chamfer rc=$?"
tail -3 gpurun_out/ncu_full_$TAG.log | cut -c1-300
