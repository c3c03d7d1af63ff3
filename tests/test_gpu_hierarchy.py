"""GPU: LOD-cut ops (gaussian_hierarchy._C surface) vs the oracle -- integer outputs and
the fp32 weights must be bit-exact (same IEEE ops, no contraction)."""
import numpy as np
import pytest

from h3dgs import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hier():
    cam = synth.make_camera(640, 360)
    leaves = synth.cloud_v1(3001, cam, zmin=2.0, zmax=30.0, seed=21)
    return cam, synth.build_hierarchy(leaves)


@pytest.mark.parametrize("tau", [0.0, 3.0, 6.0, 15.0, 200.0])
def test_expand_and_weights_bit_exact(hier, tau):
    import torch
    from gaussian_hierarchy._C import expand_to_size, get_interpolation_weights
    from oracle import oracle
    cam, h = hier
    N = h["nodes"].shape[0]
    thr = synth.tau_threshold(tau, cam)
    vp = np.array([0.3, -0.2, 1.0], np.float32)
    n_ref, ri, pi, ni = oracle.expand_to_size(h["nodes"], h["boxes"], thr, vp)
    nodes = torch.tensor(h["nodes"], device="cuda"); boxes = torch.tensor(h["boxes"], device="cuda")
    r = torch.zeros(N, dtype=torch.int32, device="cuda"); p = torch.zeros_like(r); nn_ = torch.zeros_like(r)
    n = expand_to_size(nodes, boxes, thr, torch.tensor(vp, device="cuda"), torch.zeros(3), r, p, nn_)
    assert n == n_ref and n > 0
    assert np.array_equal(r[:n].cpu().numpy(), ri) and np.array_equal(p[:n].cpu().numpy(), pi)
    assert np.array_equal(nn_[:n].cpu().numpy(), ni)
    ts_ref, kids_ref = oracle.get_interpolation_weights(ni, thr, h["nodes"], h["boxes"], vp)
    ts = torch.zeros(N, device="cuda"); kids = torch.zeros(N, dtype=torch.int32, device="cuda")
    get_interpolation_weights(nn_[:n], thr, nodes, boxes, torch.tensor(vp), torch.zeros(3), ts, kids)
    assert np.array_equal(ts[:n].cpu().numpy().view(np.uint32), ts_ref.view(np.uint32))
    assert np.array_equal(kids[:n].cpu().numpy(), kids_ref)
    if 0 < tau < 100:
        assert (ts_ref < 1).any() and (ts_ref == 1).any()


