"""`.hier` files: `load_hierarchy` / `write_hierarchy` (scene/gaussian_model.py:24, 329, 419-427;
scene/__init__.py:98-99).

LAYOUT UNPINNED.  The reader/writer of the reference live in the gaussian-hierarchy submodule, which is
absent from /root/reference, and the reference ships no `.hier` file: the byte layout below is RECALLED
from the upstream loader/writer and has not been checked against a file written by
upstream.  What IS pinned by the call sites: the argument order of both functions, the tensor shapes
(shs [P,16,3], opacity [P,1], nodes [N,7] int32, boxes [N,2,4] float32), CPU tensors out of the loader,
and that values pass through unchanged (`_scaling` is written and read back as log-scales, opacity as
the `abs`-activated value).  A round trip through this module is exact.

    int32   P                      number of Gaussians; P < 0 marks upstream's COMPRESSED variant with |P| Gaussians:
                                   positions stay float32, rotations / log-scales / opacities / SH are IEEE half
                                   (same order, same shapes) -- recalled like the rest; the exact-size check below
                                   makes a wrong recollection fail loudly instead of loading garbage
    float32 positions  [P][3]
    float32 rotations  [P][4]      w, x, y, z
    float32 log-scales [P][3]
    float32 opacities  [P]
    float32 SH         [P][16][3]  coefficient-major, as GaussianModel.get_features
    int32   N                      number of hierarchy nodes
    int32   nodes      [N][7]      depth, parent, start, count_leafs, count_merged, start_children, count_children
    float32 boxes      [N][2][4]   min.xyz + size, max.xyz + pad

Host-side disk IO, once per run: numpy, no CUDA.
"""
import os

import numpy as np
import torch

_FIELDS = (("positions", 3), ("rotations", 4), ("log_scales", 3), ("opacities", 1), ("shs", 48))


def _np(t, dtype, shape, name):
    a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
    a = np.ascontiguousarray(a, dtype=dtype)
    try:
        return a.reshape(shape)
    except ValueError:
        raise ValueError(f"write_hierarchy: {name} has shape {tuple(a.shape)}, expected {shape}") from None


def write_hierarchy(path, xyz, shs, opacities, log_scales, rotations, nodes, boxes, compressed=False):
    """Argument order of scene/gaussian_model.py:420-427.  Tensors may live on any device.  compressed (not used by the
    reference's call site): the half-precision variant, count written as -P."""
    P = int(xyz.shape[0])
    N = int(nodes.shape[0])
    ft = np.float16 if compressed else np.float32
    parts = [
        _np(xyz, np.float32, (P, 3), "xyz"), _np(rotations, np.float32, (P, 4), "rotations").astype(ft),
        _np(log_scales, np.float32, (P, 3), "scales").astype(ft), _np(opacities, np.float32, (P,), "opacities").astype(ft),
        _np(shs, np.float32, (P, 16, 3), "shs").astype(ft),
    ]
    nd, bx = _np(nodes, np.int32, (N, 7), "nodes"), _np(boxes, np.float32, (N, 2, 4), "boxes")
    tmp = f"{path}.{os.getpid()}.tmp"
    with open(tmp, "wb") as f:
        np.array([-P if compressed else P], np.int32).tofile(f)
        for a in parts:
            a.tofile(f)
        np.array([N], np.int32).tofile(f)
        nd.tofile(f)
        bx.tofile(f)
    os.replace(tmp, path)


def load_hierarchy(path):
    """-> (xyz [P,3], shs [P,16,3], opacities [P,1], log_scales [P,3], rotations [P,4], nodes [N,7] int32,
    boxes [N,2,4]) as CPU tensors, the tuple order of scene/gaussian_model.py:329."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = np.fromfile(f, np.int32, 1)
        if head.size != 1:
            raise ValueError(f"{path}: empty file")
        P = int(head[0])
        compressed = P < 0
        P = abs(P)
        width = lambda name: 4 if (name == "positions" or not compressed) else 2
        need = 4 + P * sum(w * width(name) for name, w in _FIELDS) + 4
        if size < need:
            raise ValueError(f"{path}: {size} bytes, but {P} Gaussians ({'half' if compressed else 'single'} precision) need at least {need}")
        out = {}
        for name, w in _FIELDS:
            out[name] = np.fromfile(f, np.float32 if width(name) == 4 else np.float16, P * w).astype(np.float32)
        N = int(np.fromfile(f, np.int32, 1)[0])
        if N < 0 or size != need + N * (7 * 4 + 8 * 4):
            raise ValueError(f"{path}: {size} bytes do not match P = {P}{' (compressed)' if compressed else ''}, N = {N} "
                             f"(expected {need + max(N, 0) * 60}): not the layout this module assumes")
        nodes = np.fromfile(f, np.int32, N * 7).reshape(N, 7)
        boxes = np.fromfile(f, np.float32, N * 8).reshape(N, 2, 4)
    t = torch.from_numpy
    return (t(out["positions"].reshape(P, 3)), t(out["shs"].reshape(P, 16, 3)), t(out["opacities"].reshape(P, 1)),
            t(out["log_scales"].reshape(P, 3)), t(out["rotations"].reshape(P, 4)), t(nodes), t(boxes))
