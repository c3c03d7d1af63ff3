#!/bin/bash
# One lean gpurun call (1 GPU): GPU test suite, one bench line (+ optional env-variant lines), launch list, and ONE ncu
# --set full pass that captures one launch of each of the nine hot kernels.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [full] [VAR=1 ...]'
#   full: the bench line carries the extras (value_api / value_dropin) and the CPU arms; otherwise --no-extras --no-cpu-baseline
tag=${1:-q}; shift
full=0; if [ "$1" = "full" ]; then full=1; shift; fi
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="timeout 500"
$T python -m pytest tests -m gpu -q -p no:cacheprovider -rs -x > gpurun_out/${tag}_tests.log 2>&1
tail -6 gpurun_out/${tag}_tests.log
lean="--no-extras --no-cpu-baseline"; [ $full = 1 ] && lean=""
$T python bench.py --steps 20 --warmup 3 $lean > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
if [ $full = 1 ]; then    # the record lines of a round: config #2 with the classic-formulation stand-in, the CPU reference arm, the reference-CUDA probe
  $T python bench.py --workload flat1m --classic --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_flat1m.json 2> gpurun_out/${tag}_bench_flat1m.err
  $T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
  $T python bench.py --impl reference-cuda > gpurun_out/${tag}_bench_refcuda.json 2> gpurun_out/${tag}_bench_refcuda.err
fi
for v in "$@"; do
  env $v $T python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_${v%%=*}.json 2> gpurun_out/${tag}_bench_${v%%=*}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_bench*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if "ms_per_step" not in d or "e2e" not in d or "value" not in d:
            print(f, json.dumps(d)[:300]); continue
        print(f, round(d["value"], 1), "img/s", round(d["ms_per_step"], 4), "ms  e2e", round(d["e2e"]["value"], 1), d.get("stage_ms"), "gap", d.get("host_gap_ms"),
              {k: round(d[k]["value"], 1) for k in ("value_api", "value_dropin") if k in d})
        print("   hbm_frac", d.get("stage_hbm_frac"))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-1500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --mode api --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_bench.log 2>&1
K='^(lod_cut_fused|preprocess|preprocess_color|emit_to_tiles|tile_sort_gather|render_forward|render_backward|preprocess_backward|sh_backward)_kernel'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 54 -c 9 -f -o gpurun_out/${tag}_hot \
  python bench.py --mode api --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_hot.log 2>&1
ncu -i gpurun_out/${tag}_hot.ncu-rep --page raw --csv > gpurun_out/${tag}_hot_raw.csv 2>/dev/null
python tools/ncu_split.py gpurun_out/${tag}_hot_raw.csv gpurun_out/${tag}
ls -la gpurun_out | grep "${tag}_" | head -40
