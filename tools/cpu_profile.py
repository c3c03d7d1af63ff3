"""scratch: CPU-side profile of the step (GPU box)"""
import os, sys, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hierarchical-3d-gaussians_b200"))
import torch, bench
from h3dgs import pipeline, synth
arrays, cams = bench.build_workload("hier3m")
scene = pipeline.Scene(arrays); dcams = [pipeline.DeviceCamera(c) for c in cams]
thr = [synth.tau_threshold(6.0, c) for c in cams]; bg = torch.zeros(3, device="cuda")
gts = [torch.rand((3, 1080, 1920), device="cuda") for _ in range(8)]
for i in range(16): pipeline.l1_step(scene, dcams[i % 8], bg, gts[i % 8], thr[i % 8])
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(40): pipeline.l1_step(scene, dcams[i % 8], bg, gts[i % 8], thr[i % 8])
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(32); print(s.getvalue()[:6000])
