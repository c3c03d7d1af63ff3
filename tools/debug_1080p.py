import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "hierarchical-3d-gaussians_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from h3dgs import synth
from diff_gaussian_rasterization import _C
from oracle import oracle
for (P, W, H) in [(200000, 1920, 1080), (1000000, 1920, 1080)]:
    cam = synth.make_camera(W, H)
    sc = synth.cloud_v1(P, cam)
    dev = "cuda"
    t = lambda a: torch.tensor(a, device=dev)
    m, sh, op, s, r = t(sc["means3D"]), t(sc["shs"]), t(sc["opacities"]), t(sc["scales"]), t(sc["rotations"])
    bg = torch.zeros(3, device=dev); vm = t(cam.world_view_transform); pm = t(cam.full_proj_transform); cp = t(cam.camera_center)
    e = torch.empty(0, device=dev)
    n, color, radii, gb, bb, ib, _ = _C.rasterize_gaussians(bg, m, e, op, s, r, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, sh, 3, cp, False, False, None, None, None, None, False)
    sv = _C.state_view(P, W, H, n, gb, bb, ib)
    f = oracle.rasterize_forward(sc['means3D'], sc['shs'], None, sc['opacities'], sc['scales'], sc['rotations'], None,
       cam.world_view_transform, cam.full_proj_transform, cam.camera_center, np.zeros(3,np.float32), W,H, cam.tanfovx, cam.tanfovy)
    rg = sv["ranges"].cpu().numpy().astype(np.uint32)
    print(P, "D", n, f["num_rendered"], "ranges equal", np.array_equal(rg, f["ranges"]), "len mean", (rg[:,1]-rg[:,0]).mean(), "max", (rg[:,1]-rg[:,0]).max())
    print("  color maxerr", np.abs(color.cpu().numpy()-f["color"]).max(), "ncontrib mismatch", (sv["n_contrib"].cpu().numpy().astype(np.uint32)!=f["n_contrib"]).mean())
    torch.cuda.synchronize()
    for _ in range(3):
        _C.rasterize_gaussians(bg, m, e, op, s, r, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, sh, 3, cp, False, False, None, None, None, None, False)
    torch.cuda.synchronize(); t0=time.time()
    for _ in range(5):
        _C.rasterize_gaussians(bg, m, e, op, s, r, 1.0, e, vm, pm, cam.tanfovx, cam.tanfovy, H, W, sh, 3, cp, False, False, None, None, None, None, False)
    torch.cuda.synchronize(); print("  fwd ms", (time.time()-t0)/5*1000)
