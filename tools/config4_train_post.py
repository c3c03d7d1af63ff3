#!/usr/bin/env python
"""BASELINE.json configs[3] on one B200: train_post.py post-optimisation of a 2-chunk synthetic hierarchy, 1500 views
at 1920x1080 (750 per chunk, as scripts/full_train.py runs train_post.py once per chunk).

The reference checkout does not exist on the GPU box, so the loop here is THIS repo's restatement of
/root/reference/train_post.py:70-192 -- per view: LOD threshold drawn log-uniformly in [0.005, 0.1] (:73-74),
expand_to_size + get_interpolation_weights (:91-113), render (:119-129), loss = 0.8 L1 + 0.2 (1 - SSIM) (:134-140),
backward, Adam over all six parameter groups (:167-192; lr as arguments/__init__.py:84-98).  (The reference's OWN loop
text runs on the packages under the emulation build: tests/test_gpu_reference_train_post.py.)  Two host forms:

  dropin : what the unmodified script sees -- the PyTorch gather / lerp of render_post around the rasterizer
           (h3dgs.pipeline.render_hier), the reference's conv2d SSIM formulation, dense torch.optim.Adam;
  optin  : the same iteration with the repo's opt-ins -- cut gather / lerp fused into K1/K9 (render_indices),
           fused L1 + SSIM kernel (h3dgs.loss), sparse Adam over the rows that received gradients (h3dgs.optim).

Prints one JSON line: whole-iteration milliseconds (CUDA events around the whole loop, per chunk) for both forms."""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "hierarchical-3d-gaussians_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from h3dgs import pipeline, synth  # noqa: E402
from h3dgs.loss import fused_l1_ssim  # noqa: E402
from h3dgs.optim import Adam as SparseAdam  # noqa: E402

W, H = 1920, 1080
VIEWS_PER_CHUNK = int(os.environ.get("CONFIG4_VIEWS", "750"))
LEAVES = int(os.environ.get("CONFIG4_LEAVES", "750000"))
LAMBDA = 0.2
LR = dict(means3D=0.00002, shs=0.0025, opacities=0.05, scales=0.005, rotations=0.001)


def conv_ssim(img1, img2):
    """the reference's formulation (utils/loss_utils.py:21-63): 11x11 Gaussian window, sigma 1.5, 5 grouped conv2d"""
    g = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)], device=img1.device)
    g = (g / g.sum()).unsqueeze(1)
    win = (g @ g.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous()
    a, b = img1.unsqueeze(0), img2.unsqueeze(0)
    mu1, mu2 = F.conv2d(a, win, padding=5, groups=3), F.conv2d(b, win, padding=5, groups=3)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = F.conv2d(a * a, win, padding=5, groups=3) - mu1_sq
    s2 = F.conv2d(b * b, win, padding=5, groups=3) - mu2_sq
    s12 = F.conv2d(a * b, win, padding=5, groups=3) - mu12
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu12 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def make_chunk(offset_x, seed):
    cam = synth.make_camera(W, H)
    leaves = synth.cloud_v1(LEAVES, cam, sh_degree=3, zmin=2.0, zmax=60.0, seed=seed, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    g = np.random.default_rng(7 + seed)
    leaves["scales"] = (2.4e-3 * np.sqrt(2.0 * z) * np.exp(0.5 * g.standard_normal((z.shape[0], 3)))).astype(np.float32)
    leaves["means3D"][:, 0] += offset_x
    h = synth.build_hierarchy(leaves)
    # 30 x 25 lattice walk in front of the chunk: sideways steps x small yaw steps
    cams = []
    for i in range(VIEWS_PER_CHUNK):
        ix, iy = i % 30, (i // 30) % 25
        c = synth.yaw_camera(W, H, -12.0 + iy, np.array([offset_x - 2.0 + 4.0 * ix / 29.0, 0.0, 0.0]))
        cams.append(c)
    return h, cams


def run(form, scene, dcams, gts, limits):
    params = dict(means3D=scene.means3D, shs=scene.shs, opacities=scene.opacities, scales=scene.scales, rotations=scene.rotations)
    groups = [{"params": [p], "lr": LR[k], "name": k} for k, p in params.items()]
    opt = SparseAdam(groups, lr=0.0, eps=1e-15) if form == "optin" else torch.optim.Adam(groups, lr=0.0, eps=1e-15)
    bg = torch.zeros(3, device="cuda")

    def iteration(i):
        cam, gt = dcams[i % len(dcams)], gts[i % len(gts)]
        render = pipeline.render_hier_fused if form == "optin" else pipeline.render_hier
        img, radii, n = render(scene, cam, bg, limits[i])
        img = img.clamp(0, 1)
        if form == "optin":
            loss = fused_l1_ssim(img, gt, LAMBDA)[0]
        else:
            loss = (1.0 - LAMBDA) * (img - gt).abs().mean() + LAMBDA * (1.0 - conv_ssim(img, gt))
        loss.backward()
        with torch.no_grad():
            if form == "optin":
                relevant = (scene.opacities.grad.flatten() != 0).nonzero().flatten()
                if relevant.numel():
                    opt.step(relevant)
            else:
                opt.step()
            opt.zero_grad(set_to_none=True)
        return n

    for i in range(10):
        iteration(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cuts = [iteration(i) for i in range(VIEWS_PER_CHUNK)]
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / VIEWS_PER_CHUNK, float(np.mean(cuts))


def main():
    assert torch.cuda.is_available()
    t0 = time.time()
    out = {"config": f"#4: 2 chunks x {LEAVES} leaves (N_all = {2 * (2 * LEAVES - 1)}), {2 * VIEWS_PER_CHUNK} views at {W}x{H}, SH-3, "
                     "LOD threshold log-uniform in [0.005, 0.1], loss 0.8 L1 + 0.2 (1 - SSIM), Adam", "ms_per_iteration": {}, "mean_cut": {}}
    rs = np.random.default_rng(0)
    limits = [float(2 ** (s * (math.log2(0.1) - math.log2(0.005)) + math.log2(0.005))) for s in rs.uniform(size=VIEWS_PER_CHUNK + 10)]
    gen = torch.Generator().manual_seed(5)
    gts = [torch.rand((3, H, W), generator=gen).cuda() for _ in range(8)]
    chunks = [make_chunk(off, c) for c, off in enumerate((-10.0, 10.0))]
    for form in ("optin", "dropin"):
        ms, cut = [], []
        for h, cams in chunks:
            scene = pipeline.Scene(h)
            dcams = [pipeline.DeviceCamera(x) for x in cams]
            m, n = run(form, scene, dcams, gts, limits)
            ms.append(m); cut.append(n)
            del scene
            torch.cuda.empty_cache()
        out["ms_per_iteration"][form] = float(np.mean(ms))
        out["mean_cut"][form] = float(np.mean(cut))
    out["iterations_per_s"] = {k: 1000.0 / v for k, v in out["ms_per_iteration"].items()}
    out["wall_s"] = time.time() - t0
    print(json.dumps(out))


if __name__ == "__main__":
    main()
