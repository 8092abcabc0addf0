#!/bin/bash
# N-GPU scaling check: the default bench line twice (run-to-run spread), per-rank step / own-kernel times.
# Usage: gpurun --gpus N -- 'bash tools/scale_diag.sh N TAG'
N=${1:-8}; TAG=${2:-x}
for name in ${3:-a b}; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_${name}_$TAG.log 2>&1; echo "$name rc=$?"
  grep "^{" gpurun_out/scale_${name}_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['config']['allreduce'][:4], d['clocks'], d['config']['rank_ms_step_and_own_kernels'])"
done
