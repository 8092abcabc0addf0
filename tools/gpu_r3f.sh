#!/bin/bash
TAG=${1:-r03l}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chamfer_prune.py -q -x 2>&1 | tail -3
for w in c2 c5; do timeout 300 python tools/prep_sweep.py $w probe 2>&1 | tail -3; done
timeout 300 python bench.py --configs c5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c2_$TAG.log; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/launch_summary.py gpurun_out/launches_c2_$TAG.csv | head -8
