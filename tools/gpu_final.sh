#!/bin/bash
# Last call of a round on a short budget: the full bench line first, then the GPU tests, then config #2.
tag=${1:-fin}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench.json
timeout 200 python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/${tag}_tests.log 2>&1
tail -3 gpurun_out/${tag}_tests.log
timeout 150 python bench.py --workload flat1m --classic --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_flat1m.json 2> gpurun_out/${tag}_bench_flat1m.err
tail -c 300 gpurun_out/${tag}_bench_flat1m.json
