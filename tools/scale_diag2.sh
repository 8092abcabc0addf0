#!/bin/bash
# all-reduce stream (step stream / dedicated stream) x clock sampling on / off, per-rank times.  Usage: gpurun --gpus N -- 'bash tools/scale_diag2.sh N TAG'
N=${1:-2}; TAG=${2:-x}
run() { name=$1; shift; timeout 200 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_${name}_$TAG.log 2>&1; echo "$name rc=$?";
  grep "^{" gpurun_out/scale_${name}_$TAG.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['config']['allreduce'][:4], d['config']['allreduce_stream'], d['clocks']['samples'], d['config']['rank_ms_step_and_own_kernels'])"; }
run inline_clk VPN_ALLREDUCE=nvls
run side_clk VPN_ALLREDUCE=nvls VPN_BENCH_AR_STREAM=side
run inline_noclk VPN_ALLREDUCE=nvls VPN_BENCH_NO_CLOCKS=1
