#!/bin/bash
# Capture ONE launch of each hot kernel with `ncu --set full` during a short bench run (one GPU; ncu replays
# each kernel ~40 times) and write the metric tables next to the reports.  Usage (inside a gpurun command):
#   bash tools/ncu_capture.sh [tag]        -> gpurun_out/<tag>_<kernel>.ncu-rep, gpurun_out/<tag>_<kernel>_raw.csv
# Then, back here:  python tools/ncu_summary.py gpurun_out/<tag>_<kernel>_raw.csv profiles/<name>_ncu_summary.csv "<note>"
tag=${1:-r02}
mkdir -p gpurun_out
for k in render_backward_kernel render_forward_kernel tile_sort_gather_kernel preprocess_backward_kernel sh_backward_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o gpurun_out/${tag}_$k \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_$k.log 2>&1
  ncu -i gpurun_out/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep "$tag"
