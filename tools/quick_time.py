"""scratch: quick timing of the stages on the GPU box (not the bench)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hierarchical-3d-gaussians_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from h3dgs import synth
from diff_gaussian_rasterization import _C
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
W, H = 1920, 1080
cam = synth.make_camera(W, H)
sc = synth.cloud_v1(P, cam)
dev = "cuda"
t = lambda a: torch.tensor(a, device=dev)
m, sh, op, s, r = t(sc["means3D"]), t(sc["shs"]), t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"])
bg = torch.zeros(3, device=dev); vm = t(cam.world_view_transform); pm = t(cam.full_proj_transform); cp = t(cam.camera_center)
e = torch.empty(0, device=dev)
def fwd():
    return _C.rasterize_gaussians(bg, m, e, op, s, r, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, sh, 3, cp, False, False, None, None, None, None, False)
n, color, radii, gb, bb, ib, _ = fwd()
print("P", P, "V", int((radii > 0).sum()), "D", n, "color mean", float(color.mean()))
g = torch.sign(color - torch.rand_like(color)) / color.numel()
def bwd():
    return _C.rasterize_gaussians_backward(bg, m, radii, e, op, s, r, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, g, None, sh, 3, cp, gb, n, bb, ib, False, None, None, None, None, False, H, W)
bwd()
for name, fn in [("fwd", fwd), ("bwd", bwd)]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(True), torch.cuda.Event(True)
    ev0.record()
    for _ in range(10): fn()
    ev1.record(); torch.cuda.synchronize()
    print(name, "ms", ev0.elapsed_time(ev1) / 10)
