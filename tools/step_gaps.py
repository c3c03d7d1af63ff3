"""scratch: where does the non-kernel time of a step go? (GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hierarchical-3d-gaussians_b200"))
import numpy as np, torch
import bench
from h3dgs import pipeline, synth, _lib
arrays, cams = bench.build_workload("hier3m")
dev = "cuda"
scene = pipeline.Scene(arrays, device=dev)
dcams = [pipeline.DeviceCamera(c, device=dev) for c in cams]
thr = [synth.tau_threshold(6.0, c) for c in cams]
bg = torch.zeros(3, device=dev)
gts = [torch.rand((3, 1080, 1920), device=dev) for _ in range(8)]
def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e
for i in range(5): pipeline.l1_step(scene, dcams[i % 8], bg, gts[i % 8], thr[i % 8])
torch.cuda.synchronize()
acc = np.zeros(5); cpu = np.zeros(5); K = 20
for i in range(K):
    v = i % 8
    t0 = time.perf_counter(); e0 = ev()
    scene.zero_grad()
    n = pipeline.lod_cut(scene, dcams[v], thr[v]); t1 = time.perf_counter(); e1 = ev()
    rs = pipeline.make_settings(scene, dcams[v], bg, 3, ts=scene.interpolation_weights, kids=scene.num_siblings, ridx=scene.render_indices[:n], pidx=scene.parent_indices[:n])
    m2 = torch.zeros((n, 3), device=dev, requires_grad=True)
    img, radii, _ = pipeline.GaussianRasterizer(rs)(means3D=scene.means3D, means2D=m2, shs=scene.shs, colors_precomp=None, opacities=scene.opacities, scales=scene.scales, rotations=scene.rotations, cov3D_precomp=None)
    t2 = time.perf_counter(); e2 = ev()
    loss = (img - gts[v]).abs().mean(); t3 = time.perf_counter(); e3 = ev()
    loss.backward(); t4 = time.perf_counter(); e4 = ev()
    torch.cuda.synchronize(); t5 = time.perf_counter()
    acc += [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4), e0.elapsed_time(e4)]
    cpu += [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t0]
print("GPU ms  lod/fwd/loss/bwd/total", np.round(acc / K, 3))
print("CPU ms  lod/fwd/loss/bwd/total", np.round(cpu / K * 1e3, 3))
