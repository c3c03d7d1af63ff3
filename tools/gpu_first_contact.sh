#!/bin/bash
# One gpurun call for the first contact of this branch with a B200 (every piece of it is so far checked on
# the CPU only): the GPU test suite without -x (all failures at once), the bench in its four variants, and a
# launch list.  Everything lands in gpurun_out/.  Usage:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_first_contact.sh'
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 400"
$T python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/fc_tests.log 2>&1
tail -25 gpurun_out/fc_tests.log
for v in 0 1; do
  H3DGS_GROUPWALK=$v $T python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/fc_bench_exact_gw$v.json 2> gpurun_out/fc_bench_exact_gw$v.err
  H3DGS_GROUPWALK=$v $T python bench.py --graph --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/fc_bench_graph_gw$v.json 2> gpurun_out/fc_bench_graph_gw$v.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/fc_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "img/s", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("stage_ms"))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-600:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fc_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/fc_ncu_bench.log 2>&1
bash tools/ncu_capture.sh fc > gpurun_out/fc_ncu_capture.log 2>&1
