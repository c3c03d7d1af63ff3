"""The synthetic hierarchy of h3dgs.synth.build_hierarchy, in torch ops so that large workloads
(config #5: 10 M leaves, 20 M nodes) are generated on the GPU in seconds instead of minutes of numpy on
the host.  Same algorithm and conventions (complete binary tree over Morton-sorted leaves, BFS numbering,
moment-matched merges, Node = 7 x int32, Box = 2 x float4 with the size in min.w);
tests/test_synth_torch_cpu.py checks it against the numpy version on the CPU.  Benchmark data only."""
import math

import torch

SH_C0 = 0.28209479177387814


def cloud(P, tanfovx, tanfovy, sh_degree=3, zmin=2.0, zmax=60.0, seed=0, spread=1.15, device="cpu"):
    """Cloud v2 of bench.py (hier3m) with torch's generator: world size grows as sqrt(z)."""
    g = torch.Generator(device=device).manual_seed(seed)
    u = lambda *s: torch.rand(*s, generator=g, device=device, dtype=torch.float64)
    n = lambda *s: torch.randn(*s, generator=g, device=device, dtype=torch.float64)
    z = zmin + (zmax - zmin) * u(P)
    x = z * tanfovx * (2 * spread * u(P) - spread)
    y = z * tanfovy * (2 * spread * u(P) - spread)
    means = torch.stack([x, y, z], 1).float()
    scales = (2.4e-3 * torch.sqrt(2.0 * z)[:, None] * torch.exp(0.5 * n(P, 3))).float()
    q = n(P, 4)
    rots = (q / q.norm(dim=1, keepdim=True)).float()
    opac = torch.sigmoid(1.5 * n(P)).float()[:, None]
    K = (sh_degree + 1) ** 2
    shs = torch.zeros((P, K, 3), device=device)
    shs[:, 0] = ((u(P, 3) - 0.5) / SH_C0).float()
    if K > 1:
        shs[:, 1:] = (0.1 * n(P, K - 1, 3)).float()
    return dict(means3D=means, scales=scales, rotations=rots, opacities=opac, shs=shs)


def _morton(means):
    lo, hi = means.min(0).values, means.max(0).values
    q = ((means - lo) / (hi - lo + 1e-9) * 1023).to(torch.int64)

    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)


def _R_from_quat(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.empty((q.shape[0], 3, 3), dtype=q.dtype, device=q.device)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def _quat_from_R(R):
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    w = torch.sqrt(torch.clamp(1 + m00 + m11 + m22, min=0)) / 2
    x = torch.sqrt(torch.clamp(1 + m00 - m11 - m22, min=0)) / 2
    y = torch.sqrt(torch.clamp(1 - m00 + m11 - m22, min=0)) / 2
    z = torch.sqrt(torch.clamp(1 - m00 - m11 + m22, min=0)) / 2
    x = torch.copysign(x, R[:, 2, 1] - R[:, 1, 2])
    y = torch.copysign(y, R[:, 0, 2] - R[:, 2, 0])
    z = torch.copysign(z, R[:, 1, 0] - R[:, 0, 1])
    q = torch.stack([w, x, y, z], 1)
    return q / q.norm(dim=1, keepdim=True)


def build_hierarchy(leaves):
    """leaves: dict of tensors (any device).  Returns a dict of tensors on the same device: the all-Gaussian
    arrays [N_all, ...] (float32), nodes [N_all, 7] int32, boxes [N_all, 2, 4] float32; N_all = 2 L - 1."""
    dev = leaves["means3D"].device
    L = leaves["means3D"].shape[0]
    order = torch.argsort(_morton(leaves["means3D"].double()), stable=True)
    lv = {k: v[order] for k, v in leaves.items()}
    N = 2 * L - 1
    i64 = lambda *s: torch.zeros(*s, dtype=torch.int64, device=dev)
    lo, hi, first_child, nchild = i64(N), i64(N), i64(N), i64(N)
    parent = torch.full((N,), -1, dtype=torch.int64, device=dev)
    hi[0] = L
    level = torch.zeros(1, dtype=torch.int64, device=dev)
    nxt, levels = 1, [level]
    while level.numel():
        span = hi[level] - lo[level]
        inner = level[span > 1]
        if inner.numel() == 0:
            break
        mid = (lo[inner] + hi[inner]) // 2
        c0 = nxt + 2 * torch.arange(inner.numel(), device=dev)
        c1 = c0 + 1
        lo[c0] = lo[inner]; hi[c0] = mid; lo[c1] = mid; hi[c1] = hi[inner]
        parent[c0] = inner; parent[c1] = inner; first_child[inner] = c0; nchild[inner] = 2
        nxt += 2 * inner.numel()
        level = torch.sort(torch.cat([c0, c1])).values
        levels.append(level)
    assert nxt == N
    is_leaf = (hi - lo) == 1
    K = lv["shs"].shape[1]
    f64 = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
    means, cov, opac, shs = f64(N, 3), f64(N, 3, 3), f64(N), torch.zeros((N, K, 3), dtype=torch.float32, device=dev)
    bmin, bmax, depth = f64(N, 3), f64(N, 3), i64(N)
    li = torch.nonzero(is_leaf).flatten()
    src = lo[li]
    means[li] = lv["means3D"][src].double(); opac[li] = lv["opacities"][src, 0].double(); shs[li] = lv["shs"][src]
    Rl = _R_from_quat(lv["rotations"][src].double())
    s2 = lv["scales"][src].double() ** 2
    cov[li] = torch.einsum("nik,nk,njk->nij", Rl, s2, Rl)
    ext = 3.0 * lv["scales"][src].double().max(1, keepdim=True).values
    bmin[li] = means[li] - ext; bmax[li] = means[li] + ext
    for level in reversed(levels):
        inner = level[~is_leaf[level]]
        if inner.numel() == 0:
            continue
        a, b = first_child[inner], first_child[inner] + 1
        wa = opac[a] * torch.sqrt(torch.abs(torch.linalg.det(cov[a]))) + 1e-30
        wb = opac[b] * torch.sqrt(torch.abs(torch.linalg.det(cov[b]))) + 1e-30
        ws = wa + wb
        wa = wa / ws; wb = wb / ws
        mu = wa[:, None] * means[a] + wb[:, None] * means[b]
        da, db = means[a] - mu, means[b] - mu
        cov[inner] = wa[:, None, None] * (cov[a] + da[:, :, None] * da[:, None, :]) + \
            wb[:, None, None] * (cov[b] + db[:, :, None] * db[:, None, :])
        means[inner] = mu
        opac[inner] = torch.clamp(1.15 * (wa * opac[a] + wb * opac[b]), max=1.5)    # merged weights may exceed 1
        shs[inner] = (wa[:, None, None] * shs[a].double() + wb[:, None, None] * shs[b].double()).float()
        bmin[inner] = torch.minimum(bmin[a], bmin[b]); bmax[inner] = torch.maximum(bmax[a], bmax[b])
        depth[inner] = 1 + torch.maximum(depth[a], depth[b])
    evals, evecs = torch.linalg.eigh(cov)
    flip = torch.linalg.det(evecs) < 0
    evecs[flip, :, 0] *= -1
    scales = torch.sqrt(torch.clamp(evals, min=1e-12))
    rots = _quat_from_R(evecs)
    scales[li] = lv["scales"][src].double(); rots[li] = lv["rotations"][src].double()     # leaves keep their exact values
    nodes = torch.zeros((N, 7), dtype=torch.int32, device=dev)
    nodes[:, 0] = depth.int(); nodes[:, 1] = parent.int(); nodes[:, 2] = torch.arange(N, device=dev, dtype=torch.int32)
    nodes[:, 3] = is_leaf.int(); nodes[:, 4] = (~is_leaf).int(); nodes[:, 5] = first_child.int(); nodes[:, 6] = nchild.int()
    boxes = torch.zeros((N, 2, 4), dtype=torch.float32, device=dev)
    boxes[:, 0, :3] = bmin.float(); boxes[:, 1, :3] = bmax.float()
    boxes[:, 0, 3] = (bmax - bmin).max(1).values.float()
    return dict(means3D=means.float(), scales=scales.float(), rotations=rots.float(), opacities=opac.float()[:, None],
                shs=shs, nodes=nodes, boxes=boxes)
