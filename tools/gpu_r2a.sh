#!/bin/bash
# Round-2 first GPU pass: all parity tests (incl. the reference-kernel EMD pin and BASELINE-size render tests), smoke,
# bench lines, and the ncu --set full captures the round-1 verdict asked for (pose_fwd at C2 / C3, sil_* kernels).
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_all_$TAG.log 2>&1; echo "pytest(all) rc=$?" | tee -a gpurun_out/pytest_all_$TAG.log
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 rc=$?"
for w in c3 c5; do python bench.py --workload $w --no-cpu-baseline --steps 10 > gpurun_out/bench_${w}_$TAG.log 2>&1; echo "bench $w rc=$?"; done
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pose_fwd_kernel|pose_bwd_partial|chamfer_bwd|chamfer_loss_bwd' -s 8 -c 6 -f -o gpurun_out/prof_c2step_$TAG \
    python bench.py --workload c2 --steps 2 --warmup 1 --no-cpu-baseline --graph off > gpurun_out/ncu_c2step_$TAG.log 2>&1
echo "ncu c2 step rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pose_fwd_kernel|pose_bwd_partial|sil_|chamfer_bwd|chamfer_loss_bwd' -s 16 -c 12 -f -o gpurun_out/prof_c3step_$TAG \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --graph off > gpurun_out/ncu_c3step_$TAG.log 2>&1
echo "ncu c3 step rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:sil_' -s 10 -c 5 -f -o gpurun_out/prof_c5sil_$TAG \
    python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline --graph off > gpurun_out/ncu_c5sil_$TAG.log 2>&1
echo "ncu c5 sil rc=$?"
tail -c 1500 gpurun_out/pytest_$TAG.log
