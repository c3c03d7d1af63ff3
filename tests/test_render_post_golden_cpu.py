"""CPU: the oracle's numpy restatement of render_post's gather / parent lerp / quaternion sign alignment / skybox
append (oracle/oracle.py::lerp_cut, lerp_cut_backward) against tests/golden/render_post_lerp.npz -- produced by
EXECUTING the reference's own render_post() (gaussian_renderer/__init__.py:199-234) with a capturing fake rasterizer
(tests/golden/make_golden_render_post.py).  Pins rows a14 / f-1 of SURVEY.md section 8 to the reference's code."""
import os

import numpy as np

from oracle import oracle


def load(golden_dir):
    z = np.load(os.path.join(golden_dir, "render_post_lerp.npz"))
    return {k: z[k] for k in z.files}


def cut_with_skybox(z):
    """Index form of the rows render_post hands to the rasterizer: the cut, then the skybox rows as their own parents, t = 1."""
    N, S = z["in_means3D"].shape[0], int(z["skybox_points"])
    sky = np.arange(N - S, N, dtype=np.int32)
    ri = np.concatenate([z["render_indices"], sky]); pi = np.concatenate([z["parent_indices"], sky])
    t = np.concatenate([z["t"], np.ones(S, np.float32)])
    return ri, pi, t


def test_forward_lerp_is_bit_identical_to_the_reference(golden_dir):
    z = load(golden_dir)
    ri, pi, t = cut_with_skybox(z)
    (m, sh, op, sc, rot), info = oracle.lerp_cut(z["in_means3D"], z["in_shs"], z["in_opacities"], z["in_scales"],
                                                 z["in_rotations"], ri, pi, t)
    for name, ours in (("means3D", m), ("shs", sh), ("opacities", op), ("scales", sc), ("rotations", rot)):
        assert np.array_equal(ours, z["out_" + name].reshape(ours.shape)), name
    n, S = z["render_indices"].shape[0], int(z["skybox_points"])
    # what render_post passes as weights / kids: the cut's t, then 1 / 1 on the skybox rows (:232-234)
    assert np.array_equal(z["out_interpolation_weights"][:n], z["t"]) and np.all(z["out_interpolation_weights"][n:n + S] == 1.0)
    assert np.all(z["out_num_node_kids"][n:n + S] == 1) and np.array_equal(z["out_num_node_kids"][:n], z["kids_in"][:n])
    assert int((info["sign"] < 0).sum()) > 100            # the fixture exercises the sign alignment


def test_backward_scatter_matches_the_reference_autograd(golden_dir):
    z = load(golden_dir)
    ri, pi, t = cut_with_skybox(z)
    _, info = oracle.lerp_cut(z["in_means3D"], z["in_shs"], z["in_opacities"], z["in_scales"], z["in_rotations"], ri, pi, t)
    for name in ("means3D", "shs", "opacities", "scales", "rotations"):
        g = oracle.lerp_cut_backward(z["up_" + name], info, info["sign"] if name == "rotations" else None)
        ref = z["grad_" + name]
        assert np.abs(g - ref.reshape(g.shape)).max() <= 2e-6 * np.abs(ref).max(), name      # fp32 index_add order
