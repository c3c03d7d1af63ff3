"""CPU: the torch restatement of the synthetic hierarchy builder (h3dgs.synth_torch, used to generate large
benchmark workloads on the GPU) against the numpy one (h3dgs.synth.build_hierarchy) on the same leaves."""
import numpy as np
import torch

from h3dgs import synth, synth_torch


def _cov(scales, rots):
    R = synth._R_from_quat(rots.astype(np.float64))
    return np.einsum("nik,nk,njk->nij", R, scales.astype(np.float64) ** 2, R)


def test_torch_hierarchy_matches_numpy():
    cam = synth.make_camera(640, 360)
    leaves = synth.cloud_v1(3001, cam, zmin=2.0, zmax=30.0, seed=21)
    ref = synth.build_hierarchy(leaves)
    got = synth_torch.build_hierarchy({k: torch.tensor(v) for k, v in leaves.items()})
    got = {k: v.numpy() for k, v in got.items()}
    assert np.array_equal(got["nodes"], ref["nodes"])
    for k, tol in (("boxes", 1e-6), ("means3D", 1e-6), ("opacities", 1e-6), ("shs", 1e-5)):
        assert np.allclose(got[k], ref[k], rtol=tol, atol=tol), k
    # interior scale/rotation come from an eigendecomposition (axis order and signs are a convention of the
    # LAPACK build): compare the covariances they describe
    ca, cb = _cov(got["scales"], got["rotations"]), _cov(ref["scales"], ref["rotations"])
    assert np.abs(ca - cb).max() <= 1e-5 * np.abs(cb).max()
    leaf = ref["nodes"][:, 3] == 1
    assert np.array_equal(got["scales"][leaf], ref["scales"][leaf]) and np.array_equal(got["rotations"][leaf], ref["rotations"][leaf])


def test_torch_cloud_statistics():
    cam = synth.make_camera(1920, 1080)
    c = synth_torch.cloud(20000, cam.tanfovx, cam.tanfovy, seed=3)
    z = c["means3D"][:, 2]
    assert 2.0 <= float(z.min()) and float(z.max()) <= 60.0
    assert abs(float(c["rotations"].norm(dim=1).mean()) - 1.0) < 1e-5
    assert 0.3 < float(c["opacities"].mean()) < 0.7 and c["shs"].shape == (20000, 16, 3)
    inside = (c["means3D"][:, 0].abs() <= z * cam.tanfovx) & (c["means3D"][:, 1].abs() <= z * cam.tanfovy)
    assert 0.7 < float(inside.float().mean()) < 0.8           # ~24 % frustum-culled, as cloud v1
