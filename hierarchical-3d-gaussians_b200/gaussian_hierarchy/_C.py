"""`gaussian_hierarchy._C`: LOD-cut ops on libh3dgs.so (sm_100a), same signatures
as the reference's call sites (train_post.py:91-113, render_hierarchy.py:58-80).
No CPU fallback; raises if the library is missing or a call fails.
`load_hierarchy` / `write_hierarchy` (scene/gaussian_model.py:24) are host-side disk IO: hier_io.py."""
import torch

from h3dgs import _lib
from .hier_io import load_hierarchy, write_hierarchy  # noqa: F401

_scratch = {}


def _on_device(t):
    """the library takes device pointers (the CPU suite patches this to drive an emulation build)"""
    return t.is_cuda


def _i32(t, name):
    if not _on_device(t) or t.dtype != torch.int32 or not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous int32 CUDA tensor")
    return t


def expand_to_size(nodes, boxes, size, viewpoint, viewdir, render_indices, parent_indices, nodes_for_render_indices):
    """-> int (number of Gaussians to render).  viewpoint: CUDA float tensor [3];
    viewdir: CPU float tensor [3] (unused by the size metric, kept for signature parity)."""
    L = _lib.lib()
    nodes = _i32(nodes, "nodes")
    if not _on_device(boxes) or boxes.dtype != torch.float32:
        raise RuntimeError("boxes must be a float32 CUDA tensor")
    boxes = boxes.contiguous()
    N = nodes.shape[0]
    vp = viewpoint if _on_device(viewpoint) else viewpoint.cuda()
    vp = vp.float().contiguous()
    vd = viewdir.detach().cpu().float().flatten().tolist() if viewdir is not None and viewdir.numel() >= 3 else [0.0, 0.0, 0.0]
    need = L.h3dgs_expand_scratch_bytes(N)
    key = (nodes.device.index, )
    s = _scratch.get(key)
    if s is None or s.numel() < need:
        s = torch.empty((need,), dtype=torch.uint8, device=nodes.device)
        _scratch[key] = s
    with torch.cuda.device(nodes.device):
        n = L.h3dgs_expand_to_size(N, nodes.data_ptr(), boxes.data_ptr(), float(size), vp.data_ptr(), vd[0], vd[1], vd[2],
                                   _i32(render_indices, "render_indices").data_ptr(),
                                   _i32(parent_indices, "parent_indices").data_ptr(),
                                   _i32(nodes_for_render_indices, "nodes_for_render_indices").data_ptr(),
                                   s.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return _lib.check(n)


def get_interpolation_weights(node_indices, size, nodes, boxes, viewpoint, viewdir, interpolation_weights, num_siblings):
    """Writes the first len(node_indices) entries of interpolation_weights (t) and num_siblings (k).
    viewpoint / viewdir: CPU tensors read on the host (train_post.py:109 passes camera_center.cpu())."""
    L = _lib.lib()
    n = int(node_indices.shape[0])
    if n == 0:
        return
    node_indices = _i32(node_indices.contiguous(), "node_indices")
    nodes = _i32(nodes, "nodes")
    boxes = boxes.contiguous()
    vp = viewpoint.detach().cpu().float().flatten().tolist()
    vd = viewdir.detach().cpu().float().flatten().tolist() if viewdir is not None and viewdir.numel() >= 3 else [0.0, 0.0, 0.0]
    if not _on_device(interpolation_weights) or interpolation_weights.dtype != torch.float32:
        raise RuntimeError("interpolation_weights must be a float32 CUDA tensor")
    with torch.cuda.device(nodes.device):
        _lib.check(L.h3dgs_get_interpolation_weights(n, node_indices.data_ptr(), float(size), nodes.data_ptr(),
                                                     boxes.data_ptr(), vp[0], vp[1], vp[2], vd[0], vd[1], vd[2],
                                                     interpolation_weights.data_ptr(),
                                                     _i32(num_siblings, "num_siblings").data_ptr(),
                                                     torch.cuda.current_stream().cuda_stream))
