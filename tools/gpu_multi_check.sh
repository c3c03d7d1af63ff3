#!/bin/bash
# Multi-GPU call: the peer-memory equivalence test, then the driver's launch of the sharded bench (with the N-rank check).
#   /usr/local/graft/bin/gpurun --gpus N --timeout 600 -- 'bash tools/gpu_multi_check.sh <tag> N'
tag=${1:-mN}; N=${2:-2}; shift 2
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider -rs -k peer > gpurun_out/${tag}_dist_peer_test.log 2>&1
tail -4 gpurun_out/${tag}_dist_peer_test.log
bash tools/gpu_multi_lean.sh $tag $N "$@"
