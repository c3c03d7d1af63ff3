#!/bin/bash
# Lean multi-GPU call (charged N x box time): exactly the driver's round-end launch of the sharded bench, nothing else.
#   /usr/local/graft/bin/gpurun --gpus N --timeout 600 -- 'bash tools/gpu_multi_lean.sh <tag> N [extra bench args]'
tag=${1:-mN}; N=${2:-2}; shift 2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err
tail -c 2500 gpurun_out/${tag}_bench_n$N.json; echo; tail -5 gpurun_out/${tag}_bench_n$N.err
