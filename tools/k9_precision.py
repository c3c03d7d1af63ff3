"""How much precision does the K8/K9c chain (conic -> cov2D -> cov3D -> scale/rotation) need?
numpy restatement with a dtype switch, on the test scenes; reference = the oracle (double)."""
import sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/hierarchical-3d-gaussians_b200"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from util import make_scene, oracle_run

def chain(sc, cam, cn, radii, dt, cov_dt=None):
    cov_dt = cov_dt or dt
    f = lambda a: np.asarray(a).astype(dt)
    v = f(cam.world_view_transform).reshape(-1)       # transposed storage: v[4c + k]
    m = f(sc["means3D"]); W, H = cam.W, cam.H
    fx = dt(W / (2.0 * cam.tanfovx)); fy = dt(H / (2.0 * cam.tanfovy))
    q = np.asarray(sc["rotations"]).astype(cov_dt); s = np.asarray(sc["scales"]).astype(cov_dt)
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((len(q), 3, 3), cov_dt)
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r * z); R[:, 0, 2] = 2 * (x * z + r * y)
    R[:, 1, 0] = 2 * (x * y + r * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r * x)
    R[:, 2, 0] = 2 * (x * z - r * y); R[:, 2, 1] = 2 * (y * z + r * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    Mm = s[:, :, None] * np.transpose(R, (0, 2, 1))           # Mm[k][j] = s_k R[j][k]
    V = (np.einsum("nka,nkb->nab", Mm, Mm)).astype(dt)
    tx = v[0] * m[:, 0] + v[4] * m[:, 1] + v[8] * m[:, 2] + v[12]
    ty = v[1] * m[:, 0] + v[5] * m[:, 1] + v[9] * m[:, 2] + v[13]
    tz = v[2] * m[:, 0] + v[6] * m[:, 1] + v[10] * m[:, 2] + v[14]
    limx, limy = dt(1.3 * cam.tanfovx), dt(1.3 * cam.tanfovy)
    tx = np.clip(tx / tz, -limx, limx) * tz; ty = np.clip(ty / tz, -limy, limy) * tz
    J00 = fx / tz; J02 = -(fx * tx) / (tz * tz); J11 = fy / tz; J12 = -(fy * ty) / (tz * tz)
    A = np.zeros((len(m), 2, 3), dt)
    for c in range(3):
        A[:, 0, c] = J00 * v[4 * c + 0] + J02 * v[4 * c + 2]
        A[:, 1, c] = J11 * v[4 * c + 1] + J12 * v[4 * c + 2]
    AV = np.einsum("nrk,nkc->nrc", A, V)
    a = np.einsum("nc,nc->n", AV[:, 0], A[:, 0]) + dt(0.3)
    b = np.einsum("nc,nc->n", AV[:, 0], A[:, 1])
    c_ = np.einsum("nc,nc->n", AV[:, 1], A[:, 1]) + dt(0.3)
    denom = a * c_ - b * b
    d2 = dt(1.0) / (denom * denom + dt(1e-7))
    dcx, dcy, dcz = f(cn[:, 0]), f(cn[:, 1]), f(cn[:, 2])
    dL_da = d2 * (-c_ * c_ * dcx + 2 * b * c_ * dcy + (denom - a * c_) * dcz)
    dL_dc = d2 * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c_) * dcx)
    dL_db = d2 * 2 * (b * c_ * dcx - (denom + 2 * b * b) * dcy + a * b * dcz)
    g = np.zeros((len(m), 3, 3), dt)                      # symmetric dL/dSigma with halved off-diagonals
    for i in range(3):
        for j in range(3):
            g[:, i, j] = A[:, 0, i] * A[:, 0, j] * dL_da + 0.5 * (A[:, 0, i] * A[:, 1, j] + A[:, 0, j] * A[:, 1, i]) * dL_db + A[:, 1, i] * A[:, 1, j] * dL_dc
    Mc = Mm.astype(dt); Rc = R.astype(dt); sc_ = s.astype(dt)
    dM = 2 * np.einsum("nka,naj->nkj", Mc, g)
    dscale = np.einsum("njk,nkj->nk", Rc, dM)
    dR = sc_[:, None, :] * np.transpose(dM, (0, 2, 1))    # dR[j][k] = s_k dM[k][j]
    r, x, y, z = (q[:, i].astype(dt) for i in range(4))
    dq = np.stack([
        2 * z * (dR[:, 1, 0] - dR[:, 0, 1]) + 2 * y * (dR[:, 0, 2] - dR[:, 2, 0]) + 2 * x * (dR[:, 2, 1] - dR[:, 1, 2]),
        2 * y * (dR[:, 0, 1] + dR[:, 1, 0]) + 2 * z * (dR[:, 0, 2] + dR[:, 2, 0]) + 2 * r * (dR[:, 2, 1] - dR[:, 1, 2]) - 4 * x * (dR[:, 1, 1] + dR[:, 2, 2]),
        2 * x * (dR[:, 0, 1] + dR[:, 1, 0]) + 2 * r * (dR[:, 0, 2] - dR[:, 2, 0]) + 2 * z * (dR[:, 1, 2] + dR[:, 2, 1]) - 4 * y * (dR[:, 0, 0] + dR[:, 2, 2]),
        2 * r * (dR[:, 1, 0] - dR[:, 0, 1]) + 2 * x * (dR[:, 0, 2] + dR[:, 2, 0]) + 2 * y * (dR[:, 1, 2] + dR[:, 2, 1]) - 4 * z * (dR[:, 0, 0] + dR[:, 1, 1])], 1)
    vis = radii > 0
    dscale[~vis] = 0; dq[~vis] = 0
    return dscale.astype(np.float64), dq.astype(np.float64)

rel = lambda a, b: float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
for (P, W, H, kw) in [(3000, 256, 192, dict(mode="hier", seed=42)), (20000, 640, 360, dict(seed=1)), (60000, 960, 540, dict(seed=2, scale_k=3e-3)),
                      (20000, 640, 360, dict(seed=5, scale_k=2e-2, zmax=6.0))]:
    cam, sc, ts, kids, bg = make_scene(P, W, H, **kw)
    fwd, b, gcol, gdep = oracle_run(cam, sc, bg, ts, kids)
    cn = b["conic"]
    ref_s, ref_q = b["scales"].astype(np.float64), b["rotations"].astype(np.float64)
    out = {}
    for name, dt, cdt in (("f64", np.float64, None), ("f32", np.float32, None), ("f32 chain, f64 cov", np.float32, np.float64)):
        ds, dq = chain(sc, cam, cn, fwd["radii"], dt, cdt)
        out[name] = (rel(ds, ref_s), rel(dq, ref_q))
    print(P, W, H, kw, {k: (f"{v[0]:.2e}", f"{v[1]:.2e}") for k, v in out.items()})
