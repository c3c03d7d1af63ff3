"""CPU: the hand-written forward+backward of oracle.c against an independent dense
PyTorch float64 restatement differentiated by autograd (oracle/torch_splat.py).
Two restatements that agree do not pin parity with the absent reference source,
but they do rule out chain-rule mistakes in the explicit backward."""
import numpy as np
import pytest
import torch

from oracle import oracle, torch_splat
from h3dgs import synth


def _run(mode, P=220, W=80, H=64, sh_degree=3, precomp=False):
    cam = synth.make_camera(W, H)
    sc = synth.cloud_v1(P, cam, sh_degree=sh_degree, zmin=2, zmax=8, scale_k=2e-2, seed=11)
    bg = np.array([0.2, 0.5, 0.7], np.float32)
    ts = kids = None
    if mode == "hier":
        g = np.random.default_rng(5)
        ts = g.uniform(0, 1, P).astype(np.float32); kids = g.integers(1, 5, P).astype(np.int32)
        sc["opacities"] = sc["opacities"] * 1.4          # abs-activation: opacity may exceed 1
    do_depth = mode == "depth"
    T = lambda a: None if a is None else torch.tensor(a, dtype=torch.float64)
    ins = {k: T(v).requires_grad_(True) for k, v in sc.items()}
    colors_precomp = cov_precomp = None
    if precomp:
        g = np.random.default_rng(6)
        colors_precomp = g.uniform(0, 1, (P, 3)).astype(np.float32)
        Sig = torch_splat.build_cov3d(T(sc["scales"]), T(sc["rotations"]), 1.0).numpy()
        cov_precomp = np.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], 1).astype(np.float32)
        ins["colors"] = T(colors_precomp).requires_grad_(True); ins["cov"] = T(cov_precomp).requires_grad_(True)
    f = oracle.rasterize_forward(sc["means3D"], None if precomp else sc["shs"], colors_precomp, sc["opacities"],
                                 None if precomp else sc["scales"], None if precomp else sc["rotations"], cov_precomp,
                                 cam.world_view_transform, cam.full_proj_transform, cam.camera_center, bg, W, H,
                                 cam.tanfovx, cam.tanfovy, sh_degree=sh_degree, ts=ts, kids=kids, do_depth=do_depth)
    gcol = synth.l1_grad(f["color"])
    gd = np.random.default_rng(9).standard_normal((1, H, W)).astype(np.float32) / (H * W)
    b = oracle.rasterize_backward(f, gcol, gd)
    col, radii, invd = torch_splat.splat(
        ins["means3D"], None if precomp else ins["shs"], ins.get("colors"), ins["opacities"],
        None if precomp else ins["scales"], None if precomp else ins["rotations"], ins.get("cov"),
        T(cam.world_view_transform), T(cam.full_proj_transform), T(cam.camera_center), T(bg), W, H,
        cam.tanfovx, cam.tanfovy, sh_degree=sh_degree, ts=T(ts), kids=None if kids is None else torch.tensor(kids),
        do_depth=do_depth)
    loss = (col * T(gcol)).sum() + ((invd * T(gd)).sum() if do_depth else 0)
    loss.backward()
    return f, b, col, radii, invd, ins


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(a).max(), 1e-30)


@pytest.mark.parametrize("mode", ["flat", "depth", "hier"])
def test_c_oracle_matches_torch_autograd(mode):
    f, b, col, radii, invd, ins = _run(mode)
    assert (f["radii"] > 0).sum() > 100 and f["num_rendered"] > 300
    assert (radii.numpy() == f["radii"]).all()
    assert np.abs(col.detach().numpy() - f["color"]).max() < 5e-6
    if mode == "depth":
        assert np.abs(invd.detach().numpy() - f["invdepth"]).max() < 5e-6
    tol = 1e-4 if mode == "hier" else 2e-5          # fp32 oracle vs fp64 autograd
    for k, name in [("means3D", "means3D"), ("shs", "sh"), ("opacities", "opacities"), ("scales", "scales"),
                    ("rotations", "rotations")]:
        assert _rel(ins[k].grad.numpy(), b[name]) < tol, k


def test_c_oracle_precomputed_inputs():
    f, b, col, radii, invd, ins = _run("flat", precomp=True)
    assert np.abs(col.detach().numpy() - f["color"]).max() < 5e-6
    assert _rel(ins["colors"].grad.numpy(), b["colors_precomp"]) < 2e-5
    # stored off-diagonals count twice in the symmetric matrix; autograd on the 6-vector sees the same
    assert _rel(ins["cov"].grad.numpy(), b["cov3Ds_precomp"]) < 2e-5
    assert _rel(ins["means3D"].grad.numpy(), b["means3D"]) < 2e-5


@pytest.mark.parametrize("deg", [0, 1, 2])
def test_lower_sh_degrees(deg):
    f, b, col, radii, invd, ins = _run("flat", sh_degree=deg, P=150)
    assert np.abs(col.detach().numpy() - f["color"]).max() < 5e-6
    assert _rel(ins["shs"].grad.numpy(), b["sh"]) < 2e-5
    assert _rel(ins["means3D"].grad.numpy(), b["means3D"]) < 2e-5


def test_integer_artefacts_are_consistent():
    """keys sorted, ranges partition the list by tile, n_contrib bounded by range length."""
    f, *_ = _run("flat")
    keys = f["keys"]
    assert (np.diff(keys.astype(np.uint64)) >= 0).all() if keys.size > 1 else True
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    for t in np.unique(tiles):
        s, e = f["ranges"][t]
        assert (tiles[s:e] == t).all() and (s == 0 or tiles[s - 1] != t) and (e == len(tiles) or tiles[e] != t)
    assert f["ranges"][:, 1].max() == f["num_rendered"]
    gx = (80 + 15) // 16
    ys, xs = np.mgrid[0:64, 0:80]
    tl = (ys // 16) * gx + xs // 16
    assert (f["n_contrib"] <= (f["ranges"][tl, 1] - f["ranges"][tl, 0])).all()


def test_fused_gather_lerp_oracle_matches_torch_autograd():
    """The oracle front-end's cut gather + parent lerp (+ t/(1-t) gradient scatter to the full-size arrays)
    against PyTorch ops written as in gaussian_renderer/__init__.py:199-218, differentiated by autograd."""
    cam = synth.make_camera(96, 64)
    leaves = synth.cloud_v1(160, cam, zmin=2.0, zmax=10.0, seed=4, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (3e-2 * np.sqrt(z) * np.ones((1, 3))).astype(np.float32)
    h = synth.build_hierarchy(leaves)
    thr = synth.tau_threshold(6.0, cam)
    n, ri, pi, ni = oracle.expand_to_size(h["nodes"], h["boxes"], thr, cam.camera_center)
    ts, kids = oracle.get_interpolation_weights(ni, thr, h["nodes"], h["boxes"], cam.camera_center)
    assert 0 < n < h["nodes"].shape[0] and (ts < 1).any()
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    f = oracle.rasterize_forward(h["means3D"], h["shs"], None, h["opacities"], h["scales"], h["rotations"], None,
                                 cam.world_view_transform, cam.full_proj_transform, cam.camera_center, bg, cam.W, cam.H,
                                 cam.tanfovx, cam.tanfovy, ts=ts, kids=kids, render_indices=ri, parent_indices=pi)
    gcol = synth.l1_grad(f["color"])
    b = oracle.rasterize_backward(f, gcol)
    T = lambda a: torch.tensor(a, dtype=torch.float64)
    full = {k: T(h[k]).requires_grad_(True) for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    idx = torch.tensor(ri, dtype=torch.long); par = torch.tensor(np.where(pi < 0, ri, pi), dtype=torch.long)
    t = T(ts).unsqueeze(1); u = 1 - t
    means = t * full["means3D"][idx] + u * full["means3D"][par]
    scales = t * full["scales"][idx] + u * full["scales"][par]
    shs = t.unsqueeze(2) * full["shs"][idx] + u.unsqueeze(2) * full["shs"][par]
    parents = full["rotations"][par]
    rots = full["rotations"][idx]
    dots = torch.bmm(rots.unsqueeze(1), parents.unsqueeze(2)).flatten()
    parents = torch.where((dots < 0).unsqueeze(1), -parents, parents)
    rot = t * rots + u * parents
    opac = t * full["opacities"][idx] + u * full["opacities"][par]
    col, radii, _ = torch_splat.splat(means, shs, None, opac, scales, rot, None, T(cam.world_view_transform),
                                      T(cam.full_proj_transform), T(cam.camera_center), T(bg), cam.W, cam.H, cam.tanfovx,
                                      cam.tanfovy, ts=T(ts), kids=torch.tensor(kids))
    (col * T(gcol)).sum().backward()
    assert (radii.numpy() == f["radii"]).all()
    assert np.abs(col.detach().numpy() - f["color"]).max() < 5e-6
    for k, name in [("means3D", "means3D"), ("shs", "sh"), ("opacities", "opacities"), ("scales", "scales"),
                    ("rotations", "rotations")]:
        assert _rel(full[k].grad.numpy(), b[name]) < 1e-4, k
