#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench lines, then the ncu captures (launch list + full set of the
# dominant kernels).  Usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh TAG'
TAG=${1:-r01x}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$TAG.log
python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_c2_$TAG.log 2>&1; echo "bench c2 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "bench reference rc=$?"
for w in c3 c1 c5 c4 c2f; do python bench.py --workload $w --no-cpu-baseline --steps 10 > gpurun_out/bench_${w}_$TAG.log 2>&1; echo "bench $w rc=$?"; done
python tools/bench_emd.py > gpurun_out/bench_emd_$TAG.log 2>&1; echo "bench emd rc=$?"
python tools/bench_pooling.py > gpurun_out/bench_pooling_$TAG.log 2>&1; echo "bench pooling rc=$?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
python tools/ncu_chamfer.py 0 32 > gpurun_out/plain_ncu_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:chamfer_tc_kernel|chamfer_recover' -s 3 -c 3 -f -o gpurun_out/prof_chamfer_$TAG \
    python tools/ncu_chamfer.py 0 32 > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full chamfer rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:pose_fwd_kernel|pose_bwd_partial|sil_raster|chamfer_bwd|chamfer_loss_bwd' -s 12 -c 8 -f -o gpurun_out/prof_step_$TAG \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_step_$TAG.log 2>&1
echo "ncu full step rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:emd_auction_kernel' -s 1 -c 1 -f -o gpurun_out/prof_emd_$TAG \
    python tools/bench_emd.py > gpurun_out/ncu_emd_$TAG.log 2>&1
echo "ncu full emd rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:feature_pool|image_bounds|points_yz' -s 12 -c 12 -f -o gpurun_out/prof_pool_$TAG \
    python tools/bench_pooling.py > gpurun_out/ncu_pool_$TAG.log 2>&1
echo "ncu full pooling rc=$?"
tail -c 600 gpurun_out/pytest_$TAG.log
