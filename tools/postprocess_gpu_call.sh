#!/bin/bash
# After a tools/gpu_round2.sh call: copy the evidence of <tag> from gpurun_out/ into profiles/ under <name> and print a digest.
#   bash tools/postprocess_gpu_call.sh r02a r02
tag=${1:-r02a}; name=${2:-r02}
O=gpurun_out; P=profiles
for f in bench bench_gw0 bench_k9serial bench_flat1m bench_ref bench_refcuda config4; do
  [ -s $O/${tag}_$f.json ] && tail -1 $O/${tag}_$f.json > $P/${name}_$f.json
done
[ -s $O/${tag}_launches.csv ] && cp $O/${tag}_launches.csv $P/${name}_launches_hier3m.csv
cp $O/${tag}_tests.log $P/${name}_gpu_tests.log 2>/dev/null
args=""
for kv in render_backward=render_backward_kernel render_forward=render_forward_kernel sort=tile_sort_gather_kernel \
          preprocess=preprocess_kernel key_emission=emit_to_tiles_kernel lod_cut=lod_cut_fused_kernel preprocess_color=preprocess_color_kernel \
          preprocess_backward=preprocess_backward_kernel sh_backward=sh_backward_kernel; do
  st=${kv%%=*}; k=${kv##*=}
  [ -s $O/${tag}_${k}_raw.csv ] && args="$args $st=$O/${tag}_${k}_raw.csv"
done
[ -n "$args" ] && python tools/ncu_traffic.py $tag $name $args > /dev/null
B=hierarchical-3d-gaussians_b200/build
python tools/sass_loop.py $B/render_forward.o ILb1ELb0ELb1E LDS.U8 > $P/${name}_sass_render_forward_hier_groups_loop.txt 2>/dev/null
python tools/sass_loop.py $B/render_backward.o ILb1ELb0ELb1ELb0E LDS.U8 > $P/${name}_sass_render_backward_hier_groups_loop.txt 2>/dev/null
python - <<PY
import json, glob, os
for f in sorted(glob.glob("$P/${name}_bench*.json")) + sorted(glob.glob("$P/${name}_config4.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    keep = {k: d.get(k) for k in ("value", "ms_per_step", "unavailable", "steps", "gpu_launches", "host_gap_ms", "ms_per_iteration")}
    print(os.path.basename(f), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in keep.items() if v is not None})
    for k in ("e2e", "value_api", "value_dropin", "cpu_baseline", "cpu_baseline_pytorch", "roofline", "roofline_fp32", "stage_hbm_frac", "step_roofline", "counts", "clocks", "classic_blend", "stage_ms"):
        if k in d: print("   ", k, json.dumps(d[k])[:420])
t = json.load(open("$P/ncu_traffic.json"))
for k, v in t["kernels"].items():
    print("ncu", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items() if a not in ("kernel", "source")})
PY
tail -12 $P/${name}_gpu_tests.log
