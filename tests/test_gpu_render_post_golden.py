"""GPU: both host forms of render_post's gather / parent lerp on this repo's side -- h3dgs.pipeline.interpolate_cut
(PyTorch ops, as the reference) and the fused K1/K9 form (settings.render_indices / parent_indices) -- against
tests/golden/render_post_lerp.npz, the tensors the reference's own render_post() (gaussian_renderer/__init__.py:199-234)
hands to its rasterizer and the gradients its autograd returns (tests/golden/make_golden_render_post.py)."""
import os

import numpy as np
import pytest

from h3dgs import synth

pytestmark = pytest.mark.gpu


def _scene(z):
    import torch
    from h3dgs import pipeline
    arrays = dict(means3D=z["in_means3D"], scales=z["in_scales"], rotations=z["in_rotations"], opacities=z["in_opacities"],
                  shs=z["in_shs"], skybox_points=int(z["skybox_points"]),
                  nodes=np.zeros((1, 7), np.int32), boxes=np.zeros((1, 2, 4), np.float32))
    sc = pipeline.Scene(arrays)
    n = z["render_indices"].shape[0]
    sc.render_indices[:n] = torch.tensor(z["render_indices"], device="cuda")
    sc.parent_indices[:n] = torch.tensor(z["parent_indices"], device="cuda")
    sc.interpolation_weights[:n] = torch.tensor(z["t"], device="cuda")
    sc.num_siblings[:] = torch.tensor(z["kids_in"], device="cuda")
    return sc, n


def test_pytorch_form_equals_the_reference_tensors_and_gradients(golden_dir):
    import torch
    from h3dgs import pipeline
    z = np.load(os.path.join(golden_dir, "render_post_lerp.npz"))
    sc, n = _scene(z)
    outs = dict(zip(("means3D", "scales", "rotations", "opacities", "shs"), pipeline.interpolate_cut(sc, n)))
    for k, v in outs.items():
        assert np.array_equal(v.detach().cpu().numpy(), z["out_" + k]), k          # bit-identical lerp
    sum((outs[k] * torch.tensor(z["up_" + k], device="cuda")).sum() for k in outs).backward()
    for k, p in dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities, shs=sc.shs).items():
        ref = z["grad_" + k]
        assert np.abs(p.grad.cpu().numpy() - ref).max() <= 5e-6 * np.abs(ref).max(), k      # fp32 atomic sum order


def test_fused_form_renders_the_reference_tensors_bit_identically(golden_dir):
    """K1 with render_indices/parent_indices (cut + skybox rows as indices) == K1 on the tensors the reference built;
    K9's t/(1-t) scatter == the reference's autograd applied to our per-row gradients of those tensors."""
    import torch
    from diff_gaussian_rasterization import GaussianRasterizer
    from h3dgs import pipeline
    from oracle import oracle
    z = np.load(os.path.join(golden_dir, "render_post_lerp.npz"))
    sc, n = _scene(z)
    S = sc.skybox_points
    P = n + S
    cam = synth.make_camera(96, 64)
    dcam = pipeline.DeviceCamera(cam)
    bg = torch.zeros(3, device="cuda")
    wi = torch.rand((3, cam.H, cam.W), generator=torch.Generator().manual_seed(2)).to("cuda")
    # (a) unfused: the reference's tensors straight into the rasterizer, with the weights / kids render_post passes
    t = lambda a: torch.tensor(a, device="cuda", requires_grad=True)
    ins = {k: t(z["out_" + k]) for k in ("means3D", "scales", "rotations", "opacities", "shs")}
    rs = pipeline.make_settings(sc, dcam, bg, 3, ts=torch.tensor(z["out_interpolation_weights"], device="cuda"),
                                kids=torch.tensor(z["out_num_node_kids"], device="cuda"))
    img_a, radii_a, _ = GaussianRasterizer(rs)(means3D=ins["means3D"], means2D=torch.zeros_like(ins["means3D"], requires_grad=True),
                                               shs=ins["shs"], colors_precomp=None, opacities=ins["opacities"],
                                               scales=ins["scales"], rotations=ins["rotations"], cov3D_precomp=None)
    (img_a * wi).sum().backward()
    assert int((radii_a > 0).sum()) > 50
    # (b) fused: full arrays + indices
    sky = sc.skybox_inds
    sc.render_indices[n:P] = sky; sc.parent_indices[n:P] = sky
    sc.interpolation_weights[n:P] = 1.0; sc.num_siblings[n:P] = 1
    rs = pipeline.make_settings(sc, dcam, bg, 3, ts=sc.interpolation_weights, kids=sc.num_siblings,
                                ridx=sc.render_indices[:P], pidx=sc.parent_indices[:P])
    img_b, radii_b, _ = GaussianRasterizer(rs)(means3D=sc.means3D, means2D=torch.zeros((P, 3), device="cuda", requires_grad=True),
                                               shs=sc.shs, colors_precomp=None, opacities=sc.opacities, scales=sc.scales,
                                               rotations=sc.rotations, cov3D_precomp=None)
    assert torch.equal(img_a, img_b) and torch.equal(radii_a, radii_b)
    (img_b * wi).sum().backward()
    # per-row gradients of (a) pushed through the reference-pinned scatter (tests/test_render_post_golden_cpu.py)
    sky_np = np.arange(sc.means3D.shape[0] - S, sc.means3D.shape[0], dtype=np.int32)
    ri = np.concatenate([z["render_indices"], sky_np]); pi = np.concatenate([z["parent_indices"], sky_np])
    tt = np.concatenate([z["t"], np.ones(S, np.float32)])
    _, info = oracle.lerp_cut(z["in_means3D"], z["in_shs"], z["in_opacities"], z["in_scales"], z["in_rotations"], ri, pi, tt)
    for k, p in dict(means3D=sc.means3D, scales=sc.scales, rotations=sc.rotations, opacities=sc.opacities, shs=sc.shs).items():
        ref = oracle.lerp_cut_backward(ins[k].grad.cpu().numpy(), info, info["sign"] if k == "rotations" else None)
        got = p.grad.cpu().numpy()
        assert np.abs(got - ref.reshape(got.shape)).max() <= 1e-5 * max(np.abs(ref).max(), 1e-30), k
