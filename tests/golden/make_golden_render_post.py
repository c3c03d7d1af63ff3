"""Generates tests/golden/render_post_lerp.npz by EXECUTING the reference's own render_post()
(/root/reference/gaussian_renderer/__init__.py:138-292, the interp_python=True block :199-234) on the CPU with a
CAPTURING fake rasterizer: what the reference hands to `GaussianRasterizer` after its Python gather / parent lerp /
quaternion sign alignment / skybox append, and -- through the reference's own autograd graph -- the gradients that
flow back to the full-size parameters for seeded upstream gradients.  Run in the build container only
(/root/reference does not exist on the GPU box); the fixture pins rows a14 / f-1 of SURVEY.md section 8:

  tests/test_render_post_golden_cpu.py   oracle/oracle.py's numpy restatement of the lerp + scatter  == this fixture
  tests/test_gpu_render_post_golden.py   h3dgs.pipeline.interpolate_cut (PyTorch form) and the fused K1/K9 form == this fixture
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "hierarchical-3d-gaussians_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "emul")):
    sys.path.insert(0, p)

import refharness                                    # noqa: E402
from fake_device import cuda_names_mean_cpu          # noqa: E402  (device="cuda" in the reference lands on the CPU)
from h3dgs import synth                              # noqa: E402


class Capture(torch.nn.Module):
    """Stands in for diff_gaussian_rasterization.GaussianRasterizer: records settings + arguments and returns
    differentiable dummies so that render_post() runs to its end."""
    last = None

    def __init__(self, raster_settings):
        super().__init__()
        self.rs = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None, cov3D_precomp=None):
        Capture.last = dict(rs=self.rs, means3D=means3D, means2D=means2D, opacities=opacities, shs=shs, scales=scales,
                            rotations=rotations, colors_precomp=colors_precomp, cov3D_precomp=cov3D_precomp)
        H, W = self.rs.image_height, self.rs.image_width
        return torch.zeros(3, H, W), torch.ones(means3D.shape[0], dtype=torch.int32), torch.zeros(1, H, W)


def main():
    assert refharness.have_reference()
    with cuda_names_mean_cpu():
        gr = refharness.import_reference_renderer()
        gr.GaussianRasterizer = Capture               # the module-level name render_post() looks up
        cam = synth.make_camera(96, 64)
        leaves = synth.cloud_v1(700, cam, zmin=2.0, zmax=30.0, seed=21, scale_k=1.0)
        h = synth.append_skybox(synth.build_hierarchy(leaves), 40)
        N = h["means3D"].shape[0]
        rng = np.random.default_rng(5)
        # a synthetic cut: random rows with random parents (some roots: parent -1 with t == 1, as the LOD ops emit),
        # weights in [0,1] with exact 0 / 1 mixed in; quaternions random so that about half the pairs need the sign flip
        n = 500
        ri = rng.choice(N - 40, n, replace=False).astype(np.int32)
        pi = rng.integers(0, N - 40, n).astype(np.int32)
        t = rng.uniform(0, 1, n).astype(np.float32)
        t[rng.uniform(size=n) < 0.2] = 1.0
        t[rng.uniform(size=n) < 0.05] = 0.0
        roots = rng.uniform(size=n) < 0.05
        pi[roots] = -1; t[roots] = 1.0
        h["rotations"] = (rng.standard_normal((N, 4)) * 1.0).astype(np.float32)
        h["rotations"] /= np.linalg.norm(h["rotations"], axis=1, keepdims=True)
        kids = rng.integers(1, 6, N).astype(np.int32)
        pc = refharness.StubModel(h, device="cpu")
        vcam = refharness.StubCamera(cam, device="cpu")
        # full-size scratch tensors as train_post.py:59-63; only the first n entries are meaningful
        parent_indices = torch.zeros(N, dtype=torch.int32); parent_indices[:n] = torch.tensor(pi)
        weights = torch.zeros(N); weights[:n] = torch.tensor(t)
        num_kids = torch.tensor(kids)
        gr.render_post(vcam, pc, refharness.Pipe(), torch.zeros(3), render_indices=torch.tensor(ri),
                       parent_indices=parent_indices, interpolation_weights=weights, num_node_kids=num_kids)
        c = Capture.last
        rs = c["rs"]
        assert rs.render_indices.numel() == 0 and rs.parent_indices.numel() == 0        # :244-245
        out = {k: c[k] for k in ("means3D", "opacities", "shs", "scales", "rotations")}
        P = out["means3D"].shape[0]
        assert P == n + 40 and c["means2D"].shape[0] == P
        # seeded upstream gradients on what the rasterizer received -> the reference's autograd -> full-size grads
        g = torch.Generator().manual_seed(9)
        up = {k: torch.randn(v.shape, generator=g) for k, v in out.items()}
        sum((out[k] * up[k]).sum() for k in out).backward()
        params = pc.params()
        np.savez_compressed(
            os.path.join(HERE, "render_post_lerp.npz"),
            **{f"in_{k}": v.detach().numpy() for k, v in params.items()},
            skybox_points=np.int32(40), render_indices=ri, parent_indices=pi, t=t, kids_in=kids,
            **{f"out_{k}": v.detach().numpy() for k, v in out.items()},
            out_interpolation_weights=rs.interpolation_weights.numpy(), out_num_node_kids=rs.num_node_kids.numpy(),
            **{f"up_{k}": v.numpy() for k, v in up.items()},
            **{f"grad_{k}": v.grad.numpy() for k, v in params.items()})
        print("wrote render_post_lerp.npz: n =", n, "P =", P, "sign flips =",
              int(((h["rotations"][ri] * h["rotations"][np.where(pi < 0, ri, pi)]).sum(1) < 0).sum()))


if __name__ == "__main__":
    main()
