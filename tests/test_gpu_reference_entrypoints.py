"""GPU (or the emulation build: H3DGS_EMULATE=1): the reference's OWN host code of the path runs UNMODIFIED on top
of the drop-in packages -- `gaussian_renderer.render()` (/root/reference/gaussian_renderer/__init__.py:20-136) and
`render_post()` (:138-292), driven as train_single.py:76-97 / train_post.py:91-129 drive them -- and produces
exactly what this repo's own host mirror (h3dgs.pipeline) produces from the same inputs.

The reference checkout is not vendored: these tests run where /root/reference exists (the build container, on the
emulation build through tests/test_reference_entrypoints_on_emulator_cpu.py; a GPU box that has the checkout) and
skip elsewhere."""
import math

import numpy as np
import pytest

import refharness
from h3dgs import synth
from util import make_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refharness.have_reference(), reason="reference checkout not present on this box")]


def _grads(params):
    return {k: (None if p.grad is None else p.grad.detach().cpu().numpy().copy()) for k, p in params.items()}


def _close(a, b, what):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    assert np.abs(a - b).max() <= 2e-6 * scale, (what, float(np.abs(a - b).max() / scale))   # fp32 atomic sum order only


def test_reference_render_runs_unmodified_on_the_dropin_packages():
    """train_single.py:76,97: render(viewpoint_cam, gaussians, pipe, bg) -> render / depth / viewspace_points /
    visibility_filter / radii; do_depth=True inside."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizer
    from util import cuda_settings
    gr = refharness.import_reference_renderer()
    cam, sc, _, _, bg = make_scene(6000, 320, 200, seed=11)
    pc = refharness.StubModel(sc)
    vcam = refharness.StubCamera(cam)
    bgt = torch.tensor(bg, device="cuda")
    pkg = gr.render(vcam, pc, refharness.Pipe(), bgt)
    assert set(pkg) == {"render", "depth", "viewspace_points", "visibility_filter", "radii"}
    img, depth = pkg["render"], pkg["depth"]
    assert img.shape == (3, cam.H, cam.W) and depth.shape == (1, cam.H, cam.W)
    g = torch.Generator(device="cpu").manual_seed(0)
    wi = torch.rand(img.shape, generator=g).to("cuda"); wd = torch.rand(depth.shape, generator=g).to("cuda")
    ((img * wi).sum() + (depth * wd).sum()).backward()
    ref = _grads(pc.params())
    vsp = pkg["viewspace_points"].grad.detach().cpu().numpy().copy()
    n_vis = int(pkg["visibility_filter"].numel())

    # the same call through this repo's own mirror of the reference interface
    import types
    cam2 = types.SimpleNamespace(W=cam.W, H=cam.H, world_view_transform=cam.world_view_transform,
                                 full_proj_transform=cam.full_proj_transform, camera_center=cam.camera_center,
                                 tanfovx=math.tan(vcam.FoVx * 0.5), tanfovy=math.tan(vcam.FoVy * 0.5))
    rs = cuda_settings(cam2, bg, 3, do_depth=True)
    pc2 = refharness.StubModel(sc)
    m2d = torch.zeros_like(pc2.get_xyz, requires_grad=True)
    img2, radii2, depth2 = GaussianRasterizer(raster_settings=rs)(
        means3D=pc2.get_xyz, means2D=m2d, shs=pc2.get_features, colors_precomp=None, opacities=pc2.get_opacity,
        scales=pc2.get_scaling, rotations=pc2.get_rotation, cov3D_precomp=None)
    assert torch.equal(img, img2.clamp(0, 1)) and torch.equal(depth, depth2)
    assert n_vis == int((radii2 > 0).sum().item()) and torch.equal(pkg["radii"], radii2[radii2 > 0])
    ((img2.clamp(0, 1) * wi).sum() + (depth2 * wd).sum()).backward()
    ours = _grads(pc2.params())
    for k in ref:
        _close(ref[k], ours[k], k)
    _close(vsp, m2d.grad.cpu().numpy(), "viewspace_points")
    assert np.abs(ref["means3D"]).max() > 0 and np.abs(vsp[:, :2]).max() > 0 and np.all(vsp[:, 2] == 0)


def test_reference_render_post_runs_unmodified_on_the_dropin_packages():
    """train_post.py:91-129: expand_to_size -> get_interpolation_weights -> render_post(..., render_indices=indices,
    parent_indices, interpolation_weights, num_node_kids) incl. the Python gather / parent lerp / quaternion sign flip /
    skybox rows of :199-234; compared with h3dgs.pipeline.render_hier (PyTorch-op form) and render_hier_fused (K1/K9)."""
    import torch
    from gaussian_hierarchy._C import expand_to_size, get_interpolation_weights
    from h3dgs import pipeline
    gr = refharness.import_reference_renderer()
    cam = synth.make_camera(400, 240)
    leaves = synth.cloud_v1(5000, cam, zmin=2.0, zmax=30.0, seed=7, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (5e-3 * np.sqrt(2.0 * z) * np.ones((1, 3))).astype(np.float32)
    h = synth.append_skybox(synth.build_hierarchy(leaves), 200)
    thr = synth.tau_threshold(6.0, cam)
    pc = refharness.StubModel(h)
    vcam = refharness.StubCamera(cam)
    N = pc._xyz.size(0)
    nodes, boxes = torch.tensor(h["nodes"], device="cuda"), torch.tensor(h["boxes"], device="cuda")
    # scratch exactly as train_post.py:59-63
    render_indices = torch.zeros(N).int().cuda(); parent_indices = torch.zeros(N).int().cuda()
    nodes_for_render_indices = torch.zeros(N).int().cuda()
    interpolation_weights = torch.zeros(N).float().cuda(); num_siblings = torch.zeros(N).int().cuda()
    to_render = expand_to_size(nodes, boxes, thr, vcam.camera_center, torch.zeros((3)), render_indices, parent_indices,
                               nodes_for_render_indices)
    indices = render_indices[:to_render].int()
    get_interpolation_weights(nodes_for_render_indices[:to_render], thr, nodes, boxes, vcam.camera_center.cpu(),
                              torch.zeros((3)), interpolation_weights, num_siblings)
    bgt = torch.zeros(3, device="cuda")
    pkg = gr.render_post(vcam, pc, refharness.Pipe(), bgt, render_indices=indices, parent_indices=parent_indices,
                         interpolation_weights=interpolation_weights, num_node_kids=num_siblings, use_trained_exp=True)
    assert set(pkg) == {"render", "viewspace_points", "visibility_filter", "radii"}
    img = pkg["render"]
    assert pkg["visibility_filter"].shape[0] == to_render + pc.skybox_points
    assert 0 < to_render < N and float(((interpolation_weights[:to_render] > 0) & (interpolation_weights[:to_render] < 1)).float().mean()) > 0.05
    g = torch.Generator(device="cpu").manual_seed(1)
    wi = torch.rand(img.shape, generator=g).to("cuda")
    (img * wi).sum().backward()
    ref = _grads(pc.params())

    scene = pipeline.Scene(h)
    dcam = pipeline.DeviceCamera(cam)
    dcam.tanfovx, dcam.tanfovy = math.tan(vcam.FoVx * 0.5), math.tan(vcam.FoVy * 0.5)
    for fn in (pipeline.render_hier, pipeline.render_hier_fused):
        scene.zero_grad()
        img2, radii2, n2 = fn(scene, dcam, bgt, thr)
        assert n2 == to_render
        assert torch.equal(img, img2.clamp(0, 1)), fn.__name__          # same lerp arithmetic, same kernels: bit-identical
        assert torch.equal(pkg["visibility_filter"], radii2 > 0)
        (img2.clamp(0, 1) * wi).sum().backward()
        ours = {"means3D": scene.means3D, "scales": scene.scales, "rotations": scene.rotations,
                "opacities": scene.opacities, "shs": scene.shs}
        for k in ref:
            _close(ref[k], ours[k].grad.cpu().numpy(), (fn.__name__, k))
