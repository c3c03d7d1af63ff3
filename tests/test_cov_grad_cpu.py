"""CPU: the K8 + K9c arithmetic of preprocess_backward.cu (csrc/cov_grad.cuh: fp32 with a five-value
fp64 block) compiled for the host and run on whole scenes against the double-precision oracle.
The bar is the parity bar of the GPU suite (1e-5, norm-wise); the measured error is ~1e-6."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import make_scene, oracle_run, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libcovgrad.so")
    subprocess.run(["/usr/bin/g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "emul", "cov_grad_emul.cpp"), "-lm"], check=True)
    return C.CDLL(so)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _oracle_k8_mean_part(fwd, cam, sc, d_conic, d_invd):
    """dL/dmean3D of the covariance/depth path alone: the oracle's chain rule with zero dL/dmean2D and
    precomputed colours (no SH term)."""
    from oracle import oracle
    L = oracle.lib()
    i = fwd["_in"]; P = sc["means3D"].shape[0]
    z = lambda *s: np.zeros(s, np.float32)
    d_means3D, d_cov3D, d_sh, d_scale, d_rot = z(P, 3), z(P, 6), z(P, 1, 3), z(P, 3), z(P, 4)
    zero2, zero3 = np.zeros((P, 2), np.float64), np.zeros((P, 3), np.float64)
    dummy_colors = z(P, 3)
    L.oracle_preprocess_backward(C.c_int(P), C.c_int(0), C.c_int(0), _p(i["means3D"]), _p(i["scales"]),
                                 C.c_float(i["scale_modifier"]), _p(i["rotations"]), None, None, _p(dummy_colors),
                                 _p(i["viewmatrix"]), _p(i["projmatrix"]), _p(i["campos"]), C.c_int(cam.W), C.c_int(cam.H),
                                 C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(fwd["radii"]), _p(fwd["cov3Ds"]),
                                 _p(fwd["clamped"]), _p(zero2), _p(np.ascontiguousarray(d_conic, np.float64)), _p(zero3),
                                 _p(d_invd), _p(d_means3D), _p(d_cov3D), _p(d_sh), _p(d_scale), _p(d_rot))
    return d_means3D, d_cov3D, d_scale, d_rot


@pytest.mark.parametrize("P,W,H,kw,depth", [
    (3000, 256, 192, dict(mode="hier", seed=42), False),
    (20000, 640, 360, dict(seed=1), True),
    (40000, 960, 540, dict(seed=2, scale_k=3e-3), False),
    (20000, 640, 360, dict(seed=5, scale_k=2e-2, zmax=6.0), True),          # large, close Gaussians: fov clamp active
])
def test_mixed_precision_chain_matches_the_oracle(emu, P, W, H, kw, depth):
    cam, sc, ts, kids, bg = make_scene(P, W, H, **kw)
    fwd, b, gcol, gdep = oracle_run(cam, sc, bg, ts, kids, do_depth=depth)
    d_conic = b["conic"]
    d_invd = b["invdepth"].astype(np.float64) if depth else None
    ref_mean, ref_cov, ref_scale, ref_rot = _oracle_k8_mean_part(fwd, cam, sc, d_conic, d_invd)
    assert rel_err(ref_scale, b["scales"]) < 5e-6 and rel_err(ref_rot, b["rotations"]) < 5e-6   # same chain as the full call (which takes the unrounded sums)
    f32 = lambda a: np.ascontiguousarray(a, np.float32)
    dmean, g6, dscale, dq = (np.zeros((P, k), np.float32) for k in (3, 6, 3, 4))
    emu.emu_cov_chain(C.c_int(P), _p(f32(cam.world_view_transform)), _p(f32(sc["means3D"])),
                      C.c_float(cam.W / (2.0 * cam.tanfovx)), C.c_float(cam.H / (2.0 * cam.tanfovy)),
                      C.c_float(cam.tanfovx), C.c_float(cam.tanfovy), _p(f32(sc["scales"])), C.c_float(1.0),
                      _p(f32(sc["rotations"])), None, _p(f32(d_conic)), _p(f32(d_invd)) if depth else None,
                      _p(np.ascontiguousarray(fwd["radii"], np.int32)), _p(dmean), _p(g6), _p(dscale), _p(dq))
    errs = dict(mean=rel_err(dmean, ref_mean), cov=rel_err(g6, ref_cov), scale=rel_err(dscale, ref_scale), rot=rel_err(dq, ref_rot))
    assert max(errs.values()) < 4e-6, errs
    assert (fwd["radii"] > 0).sum() > P // 4
