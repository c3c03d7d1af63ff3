"""CPU: size-independent properties of the LOD cut restated in oracle.c, on the synthetic
hierarchy of h3dgs.synth.build_hierarchy (PARITY UNPINNED: gaussian-hierarchy source absent)."""
import numpy as np
import pytest

from h3dgs import synth


@pytest.fixture(scope="module")
def hier():
    cam = synth.make_camera(640, 360)
    leaves = synth.cloud_v1(3001, cam, zmin=2.0, zmax=30.0, seed=21)
    return cam, synth.build_hierarchy(leaves)


def test_cut_is_a_valid_cut(hier):
    """size-independent property: every leaf has exactly one ancestor-or-self on the cut
    (when the viewpoint is outside every box the cut partitions the leaves)."""
    from oracle import oracle
    cam, h = hier
    vp = np.array([0.0, 0.0, -50.0], np.float32)
    n, ri, pi, ni = oracle.expand_to_size(h["nodes"], h["boxes"], synth.tau_threshold(6.0, cam), vp)
    nodes = h["nodes"]; N = nodes.shape[0]
    on_cut = np.zeros(N, bool); on_cut[ni] = True
    leaves = np.nonzero(nodes[:, 3] == 1)[0]
    cover = np.zeros(N, np.int32)
    cur = leaves.copy(); alive = np.ones(cur.size, bool)
    while alive.any():
        cover[leaves[alive & on_cut[cur]]] += 1
        nxt = nodes[cur, 1]
        alive &= nxt != -1
        cur = np.where(alive, nxt, cur)
        if not alive.any():
            break
    assert (cover[leaves] == 1).all()


def test_hierarchy_structure(hier):
    cam, h = hier
    nodes = h["nodes"]; N = nodes.shape[0]
    assert N == 2 * 3001 - 1 and nodes[0, 1] == -1 and (nodes[1:, 1] >= 0).all()
    inner = nodes[:, 6] == 2
    assert (nodes[inner, 3] == 0).all() and (nodes[inner, 4] == 1).all() and (nodes[~inner, 3] == 1).all()
    # children contiguous and pointing back
    ch = nodes[inner, 5]
    assert (nodes[ch, 1] == np.nonzero(inner)[0]).all() and (nodes[ch + 1, 1] == np.nonzero(inner)[0]).all()
    # boxes nest
    b = h["boxes"]
    p = nodes[1:, 1]
    assert (b[1:, 0, :3] >= b[p, 0, :3] - 1e-6).all() and (b[1:, 1, :3] <= b[p, 1, :3] + 1e-6).all()
    assert np.isfinite(h["scales"]).all() and (h["scales"] > 0).all()
    assert np.allclose(np.linalg.norm(h["rotations"], axis=1), 1, atol=1e-5)


def test_weights_range_and_monotone_cut(hier):
    from oracle import oracle
    cam, h = hier
    vp = np.array([0.3, -0.2, 1.0], np.float32)
    sizes = []
    for tau in [0.0, 3.0, 6.0, 15.0, 60.0]:
        thr = synth.tau_threshold(tau, cam)
        n, ri, pi, ni = oracle.expand_to_size(h["nodes"], h["boxes"], thr, vp)
        ts, kids = oracle.get_interpolation_weights(ni, thr, h["nodes"], h["boxes"], vp)
        assert ((ts >= 0) & (ts <= 1)).all() and (kids >= 1).all()
        assert (pi[ni != 0] == h["nodes"][h["nodes"][ni[ni != 0], 1], 2]).all()
        sizes.append(n)
    assert all(a >= b for a, b in zip(sizes, sizes[1:])) and sizes[0] > sizes[-1]


def test_skybox_rows_as_indices_equal_appended_rows():
    """The fused form renders the skybox by appending its row indices (own parent, t = 1) to the cut;
    render_post appends the gathered rows themselves (gaussian_renderer/__init__.py:220-234).
    Both must be the same arithmetic: 1*x + 0*x == x."""
    from oracle import oracle
    from test_gpu_pipeline import _oracle_hier_step
    cam = synth.make_camera(400, 240)
    leaves = synth.cloud_v1(5000, cam, zmin=2.0, zmax=30.0, seed=7, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (5e-3 * np.sqrt(2.0 * z) * np.ones((1, 3))).astype(np.float32)
    h = synth.append_skybox(synth.build_hierarchy(leaves), 300)
    S, N = h["skybox_points"], h["means3D"].shape[0]
    thr = synth.tau_threshold(6.0, cam)
    gt = np.random.default_rng(4).uniform(0, 1, (3, cam.H, cam.W)).astype(np.float32)
    n, f, grads = _oracle_hier_step(h, cam, thr, gt)                 # appended-rows form
    assert f["radii"].shape[0] == n + S and (f["radii"][n:] > 0).sum() > 10
    assert np.abs(grads["means3D"][-S:]).sum() > 0 and np.abs(grads["shs"][-S:]).sum() > 0
    _, ri, pi, ni = oracle.expand_to_size(h["nodes"], h["boxes"], thr, cam.camera_center)
    ts, kids = oracle.get_interpolation_weights(ni, thr, h["nodes"], h["boxes"], cam.camera_center)
    sky = np.arange(N - S, N, dtype=ri.dtype)
    g = oracle.rasterize_forward(h["means3D"], h["shs"], None, h["opacities"], h["scales"], h["rotations"], None,
                                 cam.world_view_transform, cam.full_proj_transform, cam.camera_center,
                                 np.zeros(3, np.float32), cam.W, cam.H, cam.tanfovx, cam.tanfovy,
                                 ts=np.concatenate([ts, np.ones(S, ts.dtype)]),
                                 kids=np.concatenate([kids, np.ones(S, kids.dtype)]),
                                 render_indices=np.concatenate([ri, sky]), parent_indices=np.concatenate([pi, sky]))
    assert np.array_equal(g["radii"], f["radii"])
    assert np.array_equal(g["color"], f["color"])
