#!/bin/bash
# Evidence pass: full parity suite, bench lines of every workload, ncu launch list of the default bench command, and
# ncu --set full captures of the dominant kernels (C2 Chamfer forward, C2 / C3 step kernels).
TAG=${1:-r02q}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
( time timeout 600 python bench.py ) > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 (+configs) rc=$?"
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_$TAG.log 2>&1; echo "bench reference rc=$?"
timeout 300 python bench.py --workload c1 --configs none --no-cpu-baseline > gpurun_out/bench_c1_$TAG.log 2>&1; echo "bench c1 rc=$?"
timeout 300 python tools/bench_emd.py > gpurun_out/bench_emd_$TAG.log 2>&1; echo "bench emd rc=$?"
timeout 300 python tools/bench_pooling.py > gpurun_out/bench_pooling_$TAG.log 2>&1; echo "bench pooling rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --configs none > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/ncu_chamfer.py 0 32 > gpurun_out/plain_ncu_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:chamfer_tc_kernel|chamfer_recover|chamfer_sort|chamfer_prune|chamfer_tc_plan|chamfer_row_boxes' -s 9 -c 9 -f -o gpurun_out/prof_chamfer_$TAG \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full chamfer rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pose_fwd_kernel|pose_bwd_partial|sil_|chamfer_bwd|chamfer_loss_bwd' -s 16 -c 14 -f -o gpurun_out/prof_c3step_$TAG \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_c3step_$TAG.log 2>&1
echo "ncu c3 step rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pose_fwd_kernel|pose_bwd_partial|chamfer_bwd' -s 8 -c 6 -f -o gpurun_out/prof_c2step_$TAG \
    python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --graph off --configs none > gpurun_out/ncu_c2step_$TAG.log 2>&1
echo "ncu c2 step rc=$?"
tail -c 500 gpurun_out/pytest_$TAG.log
