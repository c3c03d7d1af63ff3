#!/bin/bash
# One gpurun call: GPU test suite, the bench in its variants, launch list and ncu captures.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round2.sh <tag>'
tag=${1:-r02a}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 500"
$T python -m pytest tests -m gpu -q -p no:cacheprovider -rs > gpurun_out/${tag}_tests.log 2>&1
tail -30 gpurun_out/${tag}_tests.log
$T python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
H3DGS_GROUPWALK=0 $T python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_gw0.json 2> gpurun_out/${tag}_bench_gw0.err
H3DGS_K9_SERIAL=1 $T python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_k9serial.json 2> gpurun_out/${tag}_bench_k9serial.err
$T python bench.py --workload flat1m --classic --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_flat1m.json 2> gpurun_out/${tag}_bench_flat1m.err
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
$T python bench.py --impl reference-cuda > gpurun_out/${tag}_bench_refcuda.json 2> gpurun_out/${tag}_bench_refcuda.err
CONFIG4_VIEWS=12 CONFIG4_LEAVES=20000 $T python tools/config4_train_post.py > gpurun_out/${tag}_config4_small.json 2> gpurun_out/${tag}_config4_small.err
tail -c 600 gpurun_out/${tag}_config4_small.json; tail -3 gpurun_out/${tag}_config4_small.err
timeout 900 python tools/config4_train_post.py > gpurun_out/${tag}_config4.json 2> gpurun_out/${tag}_config4.err
tail -c 800 gpurun_out/${tag}_config4.json; tail -3 gpurun_out/${tag}_config4.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ("value", "ms_per_step", "unavailable", "steps")},
              "e2e", d.get("e2e", {}).get("value"), d.get("stage_ms"), {k: d[k]["value"] for k in ("value_api", "value_dropin") if k in d},
              d.get("cpu_baseline", {}).get("value"), d.get("classic_blend"))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-800:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --mode api --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_bench.log 2>&1
for k in render_backward_kernel render_forward_kernel tile_sort_gather_kernel preprocess_kernel emit_to_tiles_kernel lod_cut_fused_kernel preprocess_color_kernel preprocess_backward_kernel sh_backward_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$k -s 6 -c 1 -f -o gpurun_out/${tag}_$k \
    python bench.py --mode api --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_$k.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep "$tag" | head -50
