"""GPU (or the emulation build, H3DGS_EMULATE=1): BASELINE config #4 in miniature -- the reference's OWN training loop
body of train_post.py (/root/reference/train_post.py:70-192: LOD threshold sampling, expand_to_size,
get_interpolation_weights, render_post, L1 + SSIM loss, backward, anchor / skybox gradient masking, Adam step), taken
from the file's text at run time (nothing re-typed) and run for a few views on a synthetic hierarchy that is written with
write_hierarchy and loaded through the reference's GaussianModel.create_from_hier (scene/gaussian_model.py:326-399) --
all of it on top of this repo's drop-in packages.  Runs where /root/reference exists (the build container: on the
emulation build through tests/test_reference_entrypoints_on_emulator_cpu.py); skipped elsewhere."""
import json
import math
import os

import numpy as np
import pytest

import refharness
from h3dgs import synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refharness.have_reference(), reason="reference checkout not present on this box")]


class _Bar:
    def set_postfix(self, *a, **k): pass
    def update(self, *a, **k): pass
    def close(self): pass


def test_reference_train_post_loop_body_runs_on_the_dropin_packages(tmp_path):
    import torch
    from gaussian_hierarchy._C import expand_to_size, get_interpolation_weights, write_hierarchy
    gr = refharness.import_reference_renderer()
    import scene.gaussian_model as gm                      # the reference's GaussianModel (unmodified)
    from utils.loss_utils import l1_loss, ssim             # the reference's loss
    from arguments import OptimizationParams
    from argparse import ArgumentParser

    # ---- a small hierarchy on disk, in the format GaussianModel.create_from_hier reads ----
    cam = synth.make_camera(160, 112)
    leaves = synth.cloud_v1(1500, cam, zmin=2.0, zmax=30.0, seed=2, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (8e-3 * np.sqrt(2 * z) * np.ones((1, 3))).astype(np.float32)
    h = synth.build_hierarchy(leaves)
    path = str(tmp_path / "hierarchy.hier")
    t = torch.tensor
    write_hierarchy(path, t(h["means3D"]), t(h["shs"]), t(h["opacities"]), torch.log(t(h["scales"])), t(h["rotations"]),
                    t(h["nodes"]), t(h["boxes"]))
    with open(tmp_path / "exposure.json", "w") as f:
        json.dump({"synthetic": [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]}, f)
    gaussians = gm.GaussianModel(3)
    gaussians.active_sh_degree = 3
    gaussians.create_from_hier(path, 1.0, "")
    assert gaussians._xyz.shape[0] == h["means3D"].shape[0] and gaussians.nodes.shape[1] == 7
    opt = OptimizationParams(ArgumentParser())
    opt.iterations = 4
    gaussians.training_setup(opt, our_adam=False)          # train_post.py:35

    # ---- the loop body, from the reference's own text ----
    src = open(os.path.join(refharness.REF, "train_post.py")).read()
    body = src[src.index("    while iteration < opt.iterations + 1:"):src.index("def prepare_output_and_logger")]
    params = ("opt, pipe, scene, gaussians, training_generator, background, render_indices, parent_indices, "
              "nodes_for_render_indices, interpolation_weights, num_siblings, iter_start, iter_end, progress_bar, "
              "saving_iterations, checkpoint_iterations, debug_from, iteration, ema_loss_for_log, limmax, limmin")
    ns = dict(torch=torch, math=math, expand_to_size=expand_to_size, get_interpolation_weights=get_interpolation_weights,
              render_post=gr.render_post, l1_loss=l1_loss, ssim=ssim)
    exec(f"def reference_loop({params}):\n{body}", ns)

    vcam = refharness.StubCamera(cam)
    vcam.projection_matrix = vcam.full_proj_transform
    g = torch.Generator().manual_seed(3)
    vcam.original_image = torch.rand((3, cam.H, cam.W), generator=g)
    vcam.alpha_mask = None
    N = gaussians._xyz.size(0)
    z_i = lambda: torch.zeros(N).int().cuda()               # train_post.py:59-63
    before = {k: getattr(gaussians, k).detach().clone() for k in ("_xyz", "_opacity", "_scaling", "_features_dc")}
    torch.manual_seed(0)                                    # the loop draws its LOD threshold with torch.rand
    ns["reference_loop"](opt, refharness.Pipe(), None, gaussians, [[vcam]] * 4, torch.zeros(3, device="cuda"), z_i(), z_i(), z_i(),
                         torch.zeros(N).float().cuda(), z_i(), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
                         _Bar(), [], [], -1, 1, 0.0, 0.1, 0.005)
    moved = {k: float((getattr(gaussians, k).detach() - v).abs().max()) for k, v in before.items()}
    assert all(math.isfinite(v) for v in moved.values()), moved
    assert moved["_opacity"] > 0 and moved["_scaling"] > 0 and moved["_features_dc"] > 0 and moved["_xyz"] > 0, moved   # Adam stepped
    for k in ("_xyz", "_opacity", "_scaling", "_rotation", "_features_dc", "_features_rest"):
        assert bool(torch.isfinite(getattr(gaussians, k)).all()), k
