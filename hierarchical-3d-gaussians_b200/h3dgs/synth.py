"""Synthetic inputs for tests and bench (SURVEY.md section 8d "configs -> concrete
synthetic inputs").  The reference ships no generator; this is OUR spec.  numpy
only, so the same scenes are built on the CPU-only box and on the GPU box.

Camera matrices follow the reference's construction and storage exactly:
  world_view_transform = getWorld2View2(R, T).T          (scene/cameras.py:95,
  projection_matrix    = getProjectionMatrix(...).T       utils/graphics_utils.py:38-77)
  full_proj_transform  = world_view_transform @ projection_matrix   (scene/cameras.py:97)
  camera_center        = inverse(world_view_transform)[3, :3]        (scene/cameras.py:98)
(pinned against the reference's own functions in tests/golden/camera_*.npz).
"""
import math
from dataclasses import dataclass

import numpy as np

SH_C0 = 0.28209479177387814


@dataclass
class Camera:
    W: int
    H: int
    FoVx: float
    FoVy: float
    world_view_transform: np.ndarray   # [4,4] f32, transposed storage
    projection_matrix: np.ndarray      # [4,4] f32, transposed storage
    full_proj_transform: np.ndarray    # [4,4] f32
    camera_center: np.ndarray          # [3] f32

    @property
    def tanfovx(self):
        return math.tan(self.FoVx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.FoVy * 0.5)


def world2view(R, t):
    """utils/graphics_utils.py:38-50 (translate=0, scale=1)."""
    Rt = np.zeros((4, 4))
    Rt[:3, :3] = R.transpose()
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = np.linalg.inv(Rt)
    Rt = np.linalg.inv(C2W)
    return np.float32(Rt)


def projection(znear, zfar, fovX, fovY, primx=0.5, primy=0.5):
    """utils/graphics_utils.py:52-77 (z_sign = +1, off-centre principal point)."""
    tanHalfFovY = math.tan(fovY / 2)
    tanHalfFovX = math.tan(fovX / 2)
    top = tanHalfFovY * znear
    bottom = (1 - primy) * 2 * -top
    top = primy * 2 * top
    right = tanHalfFovX * znear
    left = (1 - primx) * 2 * -right
    right = primx * 2 * right
    P = np.zeros((4, 4), np.float32)
    P[0, 0] = 2.0 * znear / (right - left)
    P[1, 1] = 2.0 * znear / (top - bottom)
    P[0, 2] = (right + left) / (right - left)
    P[1, 2] = (top + bottom) / (top - bottom)
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def make_camera(W, H, fovx_deg=60.0, R=None, T=None, primx=0.5, primy=0.5, znear=0.01, zfar=100.0):
    FoVx = math.radians(fovx_deg)
    FoVy = 2.0 * math.atan(math.tan(FoVx / 2) * H / W)     # equal focal length
    R = np.eye(3) if R is None else np.asarray(R, np.float64)
    T = np.zeros(3) if T is None else np.asarray(T, np.float64)
    wv = world2view(R, T).transpose()
    pr = projection(znear, zfar, FoVx, FoVy, primx, primy).transpose()
    full = (wv[None].astype(np.float32) @ pr[None].astype(np.float32))[0]
    center = np.linalg.inv(wv)[3, :3].astype(np.float32)
    return Camera(W, H, FoVx, FoVy, np.ascontiguousarray(wv, np.float32), np.ascontiguousarray(pr, np.float32),
                  np.ascontiguousarray(full, np.float32), np.ascontiguousarray(center))


def yaw_camera(W, H, yaw_deg, trans, **kw):
    a = math.radians(yaw_deg)
    R = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    return make_camera(W, H, R=R, T=np.asarray(trans, np.float64), **kw)


def cloud_v1(P, cam, sh_degree=3, zmin=2.0, zmax=20.0, seed=0, scale_k=1.2e-3, scale_sigma=0.7,
             spread=1.15):
    """Cloud v1 of SURVEY.md 8d: ~24% frustum-culled, ~2 px median sigma at 1080p.
    Returns activated tensors as the rasterizer receives them (post-exp scales,
    normalised wxyz quaternions, opacities in (0,1), shs [P,K,3])."""
    g = np.random.default_rng(seed)
    z = g.uniform(zmin, zmax, P)
    x = z * cam.tanfovx * g.uniform(-spread, spread, P)
    y = z * cam.tanfovy * g.uniform(-spread, spread, P)
    means = np.stack([x, y, z], 1).astype(np.float32)
    logs = np.log(scale_k * z)[:, None] + scale_sigma * g.standard_normal((P, 3))
    scales = np.exp(logs).astype(np.float32)
    q = g.standard_normal((P, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    rots = q.astype(np.float32)
    g2 = np.random.default_rng(seed + 1)
    opac = (1.0 / (1.0 + np.exp(-1.5 * g2.standard_normal(P)))).astype(np.float32)[:, None]
    K = (sh_degree + 1) ** 2
    shs = np.zeros((P, K, 3), np.float32)
    shs[:, 0, :] = (g2.uniform(0, 1, (P, 3)) - 0.5) / SH_C0
    if K > 1:
        shs[:, 1:, :] = 0.1 * g2.standard_normal((P, K - 1, 3))
    return dict(means3D=means, scales=scales, rotations=rots, opacities=opac, shs=shs.astype(np.float32))


def l1_grad(image, seed=3):
    """dL/dcolor of an L1 loss against a uniform random target."""
    g = np.random.default_rng(seed)
    target = g.uniform(0, 1, image.shape).astype(np.float32)
    return (np.sign(image - target) / image.size).astype(np.float32)


# ---------------------------------------------------------------------------
# Synthetic hierarchy (config #3): complete binary tree over Morton-sorted leaves.
# Node = 7 x int32 {depth, parent, start, count_leafs, count_merged, start_children,
# count_children}; Box = 2 x float4 {min.xyz, size ; max.xyz, 0}.  Gaussian index ==
# node index (node.start = id; leaves count_leafs=1, interior count_merged=1).
# ---------------------------------------------------------------------------
def _morton(means):
    lo, hi = means.min(0), means.max(0)
    q = ((means - lo) / (hi - lo + 1e-9) * 1023).astype(np.uint64)

    def spread(v):
        v = (v | (v << 16)) & np.uint64(0x030000FF)
        v = (v | (v << 8)) & np.uint64(0x0300F00F)
        v = (v | (v << 4)) & np.uint64(0x030C30C3)
        v = (v | (v << 2)) & np.uint64(0x09249249)
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))


def _quat_from_R(R):
    """Batch rotation matrix -> wxyz quaternion (numerically safe branchless form)."""
    m00, m11, m22 = R[:, 0, 0], R[:, 1, 1], R[:, 2, 2]
    w = np.sqrt(np.maximum(0, 1 + m00 + m11 + m22)) / 2
    x = np.sqrt(np.maximum(0, 1 + m00 - m11 - m22)) / 2
    y = np.sqrt(np.maximum(0, 1 - m00 + m11 - m22)) / 2
    z = np.sqrt(np.maximum(0, 1 - m00 - m11 + m22)) / 2
    x = np.copysign(x, R[:, 2, 1] - R[:, 1, 2])
    y = np.copysign(y, R[:, 0, 2] - R[:, 2, 0])
    z = np.copysign(z, R[:, 1, 0] - R[:, 0, 1])
    q = np.stack([w, x, y, z], 1)
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _R_from_quat(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def build_hierarchy(leaves):
    """leaves: dict from cloud_v1.  Returns dict(all-Gaussian arrays [N_all,...], nodes [N_all,7] i32,
    boxes [N_all,2,4] f32).  N_all = 2L-1; node 0 is the root; children of a node are contiguous."""
    L = leaves["means3D"].shape[0]
    order = np.argsort(_morton(leaves["means3D"]), kind="stable")
    lv = {k: v[order] for k, v in leaves.items()}
    # top-down range splitting, BFS numbering
    N = 2 * L - 1
    lo = np.zeros(N, np.int64); hi = np.zeros(N, np.int64)
    parent = np.full(N, -1, np.int64); first_child = np.zeros(N, np.int64); nchild = np.zeros(N, np.int64)
    hi[0] = L
    level = np.array([0]); nxt = 1; levels = [level]
    while level.size:
        span = hi[level] - lo[level]
        inner = level[span > 1]
        if inner.size == 0:
            break
        mid = (lo[inner] + hi[inner]) // 2
        c0 = nxt + 2 * np.arange(inner.size); c1 = c0 + 1
        lo[c0] = lo[inner]; hi[c0] = mid; lo[c1] = mid; hi[c1] = hi[inner]
        parent[c0] = inner; parent[c1] = inner; first_child[inner] = c0; nchild[inner] = 2
        nxt += 2 * inner.size
        level = np.concatenate([c0, c1]); level.sort(); levels.append(level)
    assert nxt == N
    is_leaf = (hi - lo) == 1
    K = lv["shs"].shape[1]
    means = np.zeros((N, 3)); cov = np.zeros((N, 3, 3)); opac = np.zeros(N); shs = np.zeros((N, K, 3))
    bmin = np.zeros((N, 3)); bmax = np.zeros((N, 3)); depth = np.zeros(N, np.int64)
    li = np.nonzero(is_leaf)[0]; src = lo[li]
    means[li] = lv["means3D"][src]; opac[li] = lv["opacities"][src, 0]; shs[li] = lv["shs"][src]
    Rl = _R_from_quat(lv["rotations"][src].astype(np.float64)); s2 = lv["scales"][src].astype(np.float64) ** 2
    cov[li] = np.einsum("nik,nk,njk->nij", Rl, s2, Rl)
    ext = 3.0 * lv["scales"][src].max(1, keepdims=True)
    bmin[li] = means[li] - ext; bmax[li] = means[li] + ext
    for level in reversed(levels):
        inner = level[~is_leaf[level]]
        if inner.size == 0:
            continue
        a, b = first_child[inner], first_child[inner] + 1
        wa = opac[a] * np.sqrt(np.abs(np.linalg.det(cov[a]))) + 1e-30
        wb = opac[b] * np.sqrt(np.abs(np.linalg.det(cov[b]))) + 1e-30
        ws = wa + wb; wa /= ws; wb /= ws
        mu = wa[:, None] * means[a] + wb[:, None] * means[b]
        da, db = means[a] - mu, means[b] - mu
        cov[inner] = wa[:, None, None] * (cov[a] + da[:, :, None] * da[:, None, :]) + \
            wb[:, None, None] * (cov[b] + db[:, :, None] * db[:, None, :])
        means[inner] = mu
        opac[inner] = np.minimum(1.5, 1.15 * (wa * opac[a] + wb * opac[b]))   # merged weights may exceed 1
        shs[inner] = wa[:, None, None] * shs[a] + wb[:, None, None] * shs[b]
        bmin[inner] = np.minimum(bmin[a], bmin[b]); bmax[inner] = np.maximum(bmax[a], bmax[b])
        depth[inner] = 1 + np.maximum(depth[a], depth[b])
    evals, evecs = np.linalg.eigh(cov)
    flip = np.linalg.det(evecs) < 0
    evecs[flip, :, 0] *= -1
    scales = np.sqrt(np.maximum(evals, 1e-12))
    rots = _quat_from_R(evecs)
    # leaves keep their exact original scale/rotation
    scales[li] = lv["scales"][src]; rots[li] = lv["rotations"][src]
    nodes = np.zeros((N, 7), np.int32)
    nodes[:, 0] = depth; nodes[:, 1] = parent; nodes[:, 2] = np.arange(N)
    nodes[:, 3] = is_leaf; nodes[:, 4] = ~is_leaf; nodes[:, 5] = first_child; nodes[:, 6] = nchild
    boxes = np.zeros((N, 2, 4), np.float32)
    boxes[:, 0, :3] = bmin; boxes[:, 1, :3] = bmax
    boxes[:, 0, 3] = (bmax - bmin).max(1)
    return dict(means3D=means.astype(np.float32), scales=scales.astype(np.float32), rotations=rots.astype(np.float32),
                opacities=opac.astype(np.float32)[:, None], shs=shs.astype(np.float32), nodes=nodes, boxes=boxes)


def append_skybox(h, S, radius=200.0, sh_degree=3, seed=5):
    """Append S far "skybox" Gaussians after the hierarchy rows, the layout render_post expects
    (gaussian_renderer/__init__.py:220-223: the last `skybox_points` rows of the parameter tensors,
    rendered every step with t = 1, kids = 1).  nodes/boxes are unchanged (the skybox is not in the tree)."""
    g = np.random.default_rng(seed)
    d = g.standard_normal((S, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    K = h["shs"].shape[1]
    shs = np.zeros((S, K, 3), np.float32)
    shs[:, 0] = (g.uniform(0.1, 0.9, (S, 3)) - 0.5) / 0.28209479177387814
    if K > 1:
        shs[:, 1:] = 0.05 * g.standard_normal((S, K - 1, 3))
    q = g.standard_normal((S, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    sky = dict(means3D=(radius * d).astype(np.float32),
               scales=(radius * 0.03 * np.exp(0.3 * g.standard_normal((S, 3)))).astype(np.float32),
               rotations=q.astype(np.float32), opacities=g.uniform(0.3, 0.9, (S, 1)).astype(np.float32), shs=shs)
    out = dict(h)
    for k, v in sky.items():
        out[k] = np.concatenate([h[k], v])
    out["skybox_points"] = S
    return out


def tau_threshold(tau, cam):
    """render_hierarchy.py:55-56"""
    return (2 * (tau + 0.5)) * cam.tanfovx / (0.5 * cam.W)
