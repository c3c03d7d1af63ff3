"""GPU: end-to-end sanity of the whole path as a training step -- a few Adam iterations of the
hierarchy post-optimisation loop (train_post.py:91-192 in miniature: LOD cut -> fused gather/lerp ->
rasterize -> L1 -> backward -> Adam on the full-size parameters) must reduce the loss towards a
target rendered from perturbed parameters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_post_optimisation_loop_reduces_loss():
    import torch
    from h3dgs import synth, pipeline
    cam = synth.make_camera(320, 180)
    leaves = synth.cloud_v1(6000, cam, zmin=2.0, zmax=30.0, seed=5, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (8e-3 * np.sqrt(2.0 * z) * np.ones((1, 3))).astype(np.float32)
    h = synth.build_hierarchy(leaves)
    thr = synth.tau_threshold(6.0, cam)
    dcam = pipeline.DeviceCamera(cam)
    bg = torch.zeros(3, device="cuda")
    target_scene = pipeline.Scene(h, requires_grad=False)
    with torch.no_grad():
        gt = pipeline.render_hier_fused(target_scene, dcam, bg, thr)[0].clone()
    g = np.random.default_rng(0)
    h2 = dict(h)
    h2["shs"] = (h["shs"] + 0.15 * g.standard_normal(h["shs"].shape)).astype(np.float32)
    h2["opacities"] = np.clip(h["opacities"] * g.uniform(0.6, 1.0, h["opacities"].shape), 0.01, None).astype(np.float32)
    scene = pipeline.Scene(h2)
    opt = torch.optim.Adam([{"params": [scene.shs], "lr": 2e-2}, {"params": [scene.opacities], "lr": 1e-2}])
    losses = []
    for it in range(40):
        loss, radii, n = pipeline.l1_step(scene, dcam, bg, gt, thr)
        losses.append(loss.item())
        opt.step()
    assert np.isfinite(losses).all()
    assert losses[-1] < 0.5 * losses[0], (losses[0], losses[-1])
    for p in scene.params():
        assert torch.isfinite(p.grad).all()
