"""CPU: `.hier` reader/writer (gaussian_hierarchy._C.load_hierarchy / write_hierarchy).  The byte layout
is RECALLED, not pinned (no upstream file or source here): these tests fix the call-site contract
(scene/gaussian_model.py:329, 419-427), the documented layout, and exact round trips."""
import os
import struct

import numpy as np
import pytest
import torch

from h3dgs import synth


def _hier(L=257, seed=4):
    cam = synth.make_camera(320, 200)
    return synth.build_hierarchy(synth.cloud_v1(L, cam, seed=seed))


def test_round_trip_is_exact_and_matches_the_call_site_contract(tmp_path):
    from gaussian_hierarchy._C import load_hierarchy, write_hierarchy
    h = _hier()
    t = lambda k: torch.tensor(h[k])
    path = str(tmp_path / "scene.hier")
    log_scales = torch.log(t("scales"))
    # argument order of GaussianModel.save_hier (scene/gaussian_model.py:420-427)
    write_hierarchy(path, t("means3D"), t("shs"), t("opacities"), log_scales, t("rotations"), t("nodes"), t("boxes"))
    xyz, shs_all, alpha, scales, rots, nodes, boxes = load_hierarchy(path)     # tuple order of :329
    P, N = h["means3D"].shape[0], h["nodes"].shape[0]
    assert xyz.shape == (P, 3) and shs_all.shape == (P, 16, 3) and alpha.shape == (P, 1)
    assert scales.shape == (P, 3) and rots.shape == (P, 4)
    assert nodes.shape == (N, 7) and nodes.dtype == torch.int32 and boxes.shape == (N, 2, 4)
    for got, ref in ((xyz, t("means3D")), (shs_all, t("shs")), (alpha, t("opacities")), (scales, log_scales),
                     (rots, t("rotations")), (nodes, t("nodes")), (boxes, t("boxes"))):
        assert not got.is_cuda and torch.equal(got, ref)
    # what create_from_hier does with it (scene/gaussian_model.py:385-389)
    assert shs_all[:, :1, :].shape == (P, 1, 3) and shs_all[:, 1:16, :].shape == (P, 15, 3)


def test_documented_byte_layout(tmp_path):
    from gaussian_hierarchy._C import write_hierarchy
    h = _hier(L=5)
    P, N = h["means3D"].shape[0], h["nodes"].shape[0]
    path = str(tmp_path / "tiny.hier")
    write_hierarchy(path, *(torch.tensor(h[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations", "nodes", "boxes")))
    raw = open(path, "rb").read()
    assert len(raw) == 4 + P * 4 * (3 + 4 + 3 + 1 + 48) + 4 + N * (28 + 32)
    assert struct.unpack_from("<i", raw, 0)[0] == P
    o = 4
    assert np.array_equal(np.frombuffer(raw, np.float32, 3 * P, o).reshape(P, 3), h["means3D"]); o += 12 * P
    assert np.array_equal(np.frombuffer(raw, np.float32, 4 * P, o).reshape(P, 4), h["rotations"]); o += 16 * P
    assert np.array_equal(np.frombuffer(raw, np.float32, 3 * P, o).reshape(P, 3), h["scales"]); o += 12 * P
    assert np.array_equal(np.frombuffer(raw, np.float32, P, o), h["opacities"][:, 0]); o += 4 * P
    assert np.array_equal(np.frombuffer(raw, np.float32, 48 * P, o).reshape(P, 16, 3), h["shs"]); o += 192 * P
    assert struct.unpack_from("<i", raw, o)[0] == N; o += 4
    assert np.array_equal(np.frombuffer(raw, np.int32, 7 * N, o).reshape(N, 7), h["nodes"]); o += 28 * N
    assert np.array_equal(np.frombuffer(raw, np.float32, 8 * N, o).reshape(N, 2, 4), h["boxes"])


def test_compressed_variant_round_trip(tmp_path):
    """count written as -P: positions stay float32, rotations / log-scales / opacities / SH travel as IEEE half (layout
    recalled, like the rest); the loader hands back float32 tensors of the call-site shapes"""
    from gaussian_hierarchy._C import load_hierarchy, write_hierarchy
    h = _hier(L=33)
    P, N = h["means3D"].shape[0], h["nodes"].shape[0]
    path = str(tmp_path / "c.hier")
    args = [torch.tensor(h[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations", "nodes", "boxes")]
    write_hierarchy(path, *args, compressed=True)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == -P
    assert len(raw) == 4 + P * (12 + 2 * (4 + 3 + 1 + 48)) + 4 + N * 60
    xyz, shs_all, alpha, scales, rots, nodes, boxes = load_hierarchy(path)
    assert torch.equal(xyz, args[0]) and torch.equal(nodes, args[5]) and torch.equal(boxes, args[6])
    half = lambda t: t.to(torch.float16).to(torch.float32)
    assert torch.equal(shs_all, half(args[1])) and torch.equal(alpha, half(args[2]))
    assert torch.equal(scales, half(args[3])) and torch.equal(rots, half(args[4]))
    assert all(t.dtype == torch.float32 for t in (xyz, shs_all, alpha, scales, rots))


def test_malformed_files_are_rejected(tmp_path):
    from gaussian_hierarchy._C import load_hierarchy, write_hierarchy
    h = _hier(L=9)
    path = str(tmp_path / "a.hier")
    write_hierarchy(path, *(torch.tensor(h[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations", "nodes", "boxes")))
    raw = open(path, "rb").read()
    trunc = str(tmp_path / "trunc.hier")
    open(trunc, "wb").write(raw[:len(raw) - 10])
    with pytest.raises(ValueError):
        load_hierarchy(trunc)
    half = str(tmp_path / "half.hier")                          # a single-precision body under a "compressed" count: sizes disagree
    open(half, "wb").write(struct.pack("<i", -h["means3D"].shape[0]) + raw[4:])
    with pytest.raises(ValueError, match="do not match"):
        load_hierarchy(half)
    with pytest.raises(ValueError, match="shape"):
        write_hierarchy(path, torch.zeros(4, 3), torch.zeros(4, 15, 3), torch.zeros(4, 1), torch.zeros(4, 3), torch.zeros(4, 4),
                        torch.zeros(1, 7, dtype=torch.int32), torch.zeros(1, 2, 4))
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".tmp")]


def test_loaded_hierarchy_feeds_the_lod_cut_oracle(tmp_path):
    """a file written from a synthetic hierarchy and read back drives the same cut"""
    from gaussian_hierarchy._C import load_hierarchy, write_hierarchy
    from oracle import oracle
    cam = synth.make_camera(320, 200)
    h = synth.build_hierarchy(synth.cloud_v1(600, cam, seed=9))
    path = str(tmp_path / "b.hier")
    write_hierarchy(path, *(torch.tensor(h[k]) for k in ("means3D", "shs", "opacities", "scales", "rotations", "nodes", "boxes")))
    *_, nodes, boxes = load_hierarchy(path)
    thr = synth.tau_threshold(6.0, cam)
    a = oracle.expand_to_size(h["nodes"], h["boxes"], thr, cam.camera_center)
    b = oracle.expand_to_size(nodes.numpy(), boxes.numpy(), thr, cam.camera_center)
    assert a[0] == b[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], b[1:]))
