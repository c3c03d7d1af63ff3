#!/bin/bash
# Multi-GPU call: /usr/local/graft/bin/gpurun --gpus N --timeout 1200 -- 'bash tools/gpu_multi.sh <tag> N [workloads...]'
# NCCL / peer-memory equivalence tests, then the sharded bench (peer mode and, for comparison, NCCL collectives).
tag=${1:-m2}; N=${2:-2}; shift 2
WL=${@:-hier3m}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -q -p no:cacheprovider -rs > gpurun_out/${tag}_parity_tests.log 2>&1
tail -4 gpurun_out/${tag}_parity_tests.log
timeout 600 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider -rs > gpurun_out/${tag}_dist_tests.log 2>&1
tail -15 gpurun_out/${tag}_dist_tests.log
run() {  # name, extra args
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 20 --warmup 5 "$@" > gpurun_out/${tag}_${name}.json 2> gpurun_out/${tag}_${name}.err
  tail -c 1500 gpurun_out/${tag}_${name}.json | head -c 1500; echo; tail -3 gpurun_out/${tag}_${name}.err
}
# one rank alone on this box: the single-GPU line of the same build (stage times of the kernels changed since the last call)
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
    print("N=1", round(d["value"], 1), "img/s", d["stage_ms"], "gap", d.get("host_gap_ms"))
except Exception as e:
    print("N=1 bench failed", e, open("gpurun_out/${tag}_bench_n1.err").read()[-800:])
PY
for w in $WL; do
  run bench_${w}_peer --workload $w --mode graph
  if [ "$w" = "hier3m" ]; then run bench_${w}_nccl --workload $w --mode graph --no-peer --no-extras; run bench_${w}_api --workload $w --mode api --no-extras; fi
done
