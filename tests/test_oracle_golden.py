"""CPU: pin the oracle against fixtures generated from the reference's own Python
(tests/golden/make_golden.py).  These are the only parts of the path whose
definition is present in /root/reference; everything else is "parity unpinned"."""
import os

import numpy as np
import pytest

from oracle import oracle
from h3dgs import synth


def _front_camera(W=256, H=256, push=25.0):
    # camera far behind the cloud so that every golden point is in front of the near plane
    return synth.make_camera(W, H, fovx_deg=90.0, T=np.array([0.0, 0.0, push]))


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_colour_matches_reference_eval_sh(golden_dir, deg):
    z = np.load(os.path.join(golden_dir, f"sh_deg{deg}.npz"))
    P = z["xyz"].shape[0]
    cam = _front_camera()
    scales = np.full((P, 3), 0.05, np.float32)
    rots = np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1))
    f = oracle.rasterize_forward(z["xyz"], z["feats"], None, np.full((P, 1), 0.5, np.float32), scales, rots, None,
                                 cam.world_view_transform, cam.full_proj_transform, z["campos"],
                                 np.zeros(3, np.float32), cam.W, cam.H, cam.tanfovx, cam.tanfovy, sh_degree=deg)
    vis = f["radii"] > 0
    assert vis.sum() > P // 2
    np.testing.assert_allclose(f["rgb"][vis], z["colors"][vis], rtol=2e-5, atol=2e-6)
    assert (f["clamped"][vis].astype(bool) == z["clamped"][vis]).all()


def test_cov3d_matches_reference_build_scaling_rotation(golden_dir):
    z = np.load(os.path.join(golden_dir, "cov3d.npz"))
    P = z["scales"].shape[0]
    cam = _front_camera()
    g = np.random.default_rng(0)
    xyz = g.uniform(-2, 2, (P, 3)).astype(np.float32)
    f = oracle.rasterize_forward(xyz, np.zeros((P, 1, 3), np.float32), None, np.full((P, 1), 0.5, np.float32),
                                 z["scales"], z["rotations"], None, cam.world_view_transform, cam.full_proj_transform,
                                 cam.camera_center, np.zeros(3, np.float32), cam.W, cam.H, cam.tanfovx, cam.tanfovy,
                                 sh_degree=0)
    vis = f["radii"] > 0
    assert vis.sum() > P // 2
    np.testing.assert_allclose(f["cov3Ds"][vis], z["cov6"][vis], rtol=1e-5, atol=1e-6)


def test_camera_construction_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "camera.npz"))
    for i in range(6):
        W, H = 640, 480
        fovx, fovy = float(z[f"fovx_{i}"]), float(z[f"fovy_{i}"])
        wv = synth.world2view(z[f"R_{i}"], z[f"T_{i}"]).transpose()
        pr = synth.projection(0.01, 100.0, fovx, fovy, float(z[f"primx_{i}"]), float(z[f"primy_{i}"])).transpose()
        np.testing.assert_allclose(wv, z[f"wv_{i}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(pr, z[f"pr_{i}"], rtol=1e-6, atol=1e-7)
        full = wv.astype(np.float32) @ pr.astype(np.float32)
        np.testing.assert_allclose(full, z[f"full_{i}"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(np.linalg.inv(wv)[3, :3], z[f"center_{i}"], rtol=1e-5, atol=1e-6)


def test_projection_convention_point_lands_where_expected():
    """A point on the optical axis projects to the image centre; +x goes right, +y goes down
    (utils/graphics_utils.py:52-77 with z_sign=+1)."""
    cam = synth.make_camera(128, 96)
    pts = np.array([[0, 0, 5], [1, 0, 5], [0, 1, 5]], np.float32)
    P = 3
    f = oracle.rasterize_forward(pts, np.zeros((P, 1, 3), np.float32), None, np.full((P, 1), 0.5, np.float32),
                                 np.full((P, 3), 0.01, np.float32), np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1)),
                                 None, cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                 np.zeros(3, np.float32), cam.W, cam.H, cam.tanfovx, cam.tanfovy, sh_degree=0)
    np.testing.assert_allclose(f["xy"][0], [63.5, 47.5], atol=1e-3)
    assert f["xy"][1, 0] > f["xy"][0, 0] and abs(f["xy"][1, 1] - f["xy"][0, 1]) < 1e-3
    assert f["xy"][2, 1] > f["xy"][0, 1]
    np.testing.assert_allclose(f["depths"], [5, 5, 5], rtol=1e-6)
