#!/bin/bash
TAG=${1:-r04b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_train_loop_integration.py -q -x 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "silhouette or render or step or end_to_end" 2>&1 | tail -3
timeout 300 python bench.py --workload c3 --configs c5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c3_$TAG.log; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3_$TAG.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_launches_c3_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_c3_$TAG.csv 2>/dev/null | grep "sil_\|total us"
