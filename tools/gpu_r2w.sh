#!/bin/bash
TAG=${1:-r02w}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "chamfer or sampling or transform or end_to_end or step" 2>&1 | tail -2
VPN_BENCH_NO_SIL_OVERLAP=1 timeout 300 python bench.py --workload c3 --configs c5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c3_nooverlap_$TAG.log
timeout 300 python bench.py --workload c3 --configs c5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c3_overlap_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pose_bwd_partial|chamfer_bwd_rows' -s 4 -c 6 -f -o gpurun_out/prof_bwd_$TAG \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_bwd_$TAG.log 2>&1
echo "ncu rc=$?"
