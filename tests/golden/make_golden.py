"""Generates tests/golden/*.npz by IMPORTING the reference's own Python (run in the
build container only; /root/reference does not exist on the GPU box).

Pins the parts of the oracle that the reference's present files define:
  sh_deg{0..3}.npz : utils/sh_utils.py:57-112 eval_sh + the +0.5/clamp of
                     gaussian_renderer/__init__.py:85-89
  cov3d.npz        : utils/general_utils.py:68-114 build_scaling_rotation/strip_symmetric,
                     composed as scene/gaussian_model.py:30-34
  camera.npz       : utils/graphics_utils.py:38-77, composed as scene/cameras.py:95-98
The hierarchy-rasterizer / gaussian-hierarchy kernels themselves are absent from
/root/reference (empty submodules) -- no golden vectors can be produced for them.
"""
import math
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

# the reference helpers allocate with device="cuda"; run them on CPU unchanged otherwise
_zeros = torch.zeros
def _cpu_zeros(*a, **k):
    k.pop("device", None)
    return _zeros(*a, **k)
torch.zeros = _cpu_zeros

from utils.sh_utils import eval_sh                      # noqa: E402
from utils.general_utils import build_scaling_rotation, strip_symmetric   # noqa: E402
from utils.graphics_utils import getWorld2View2, getProjectionMatrix      # noqa: E402

g = torch.Generator().manual_seed(1234)
N = 257

# ---- SH colour (features [N,16,3] as scene/gaussian_model.py:121-124) ----
feats = torch.randn(N, 16, 3, generator=g) * 0.4
xyz = torch.randn(N, 3, generator=g) * 3
campos = torch.tensor([0.3, -0.2, 0.1])
for deg in range(4):
    shs_view = feats.transpose(1, 2).view(-1, 3, 16)
    dir_pp = xyz - campos.repeat(N, 1)
    dirn = dir_pp / dir_pp.norm(dim=1, keepdim=True)
    sh2rgb = eval_sh(deg, shs_view, dirn)
    col = torch.clamp_min(sh2rgb + 0.5, 0.0)
    np.savez(os.path.join(HERE, f"sh_deg{deg}.npz"), feats=feats.numpy(), xyz=xyz.numpy(), campos=campos.numpy(),
             colors=col.numpy(), clamped=(sh2rgb + 0.5 < 0).numpy())

# ---- cov3D ----
s = torch.exp(torch.randn(N, 3, generator=g))
q = torch.randn(N, 4, generator=g)
qn = torch.nn.functional.normalize(q)
for mod in (1.0,):
    L = build_scaling_rotation(mod * s, qn)
    cov = strip_symmetric(L @ L.transpose(1, 2))
np.savez(os.path.join(HERE, "cov3d.npz"), scales=s.numpy(), rotations=qn.numpy(), cov6=cov.numpy())

# ---- camera matrices ----
cams = []
rs = np.random.RandomState(7)
for i in range(6):
    a, b = rs.uniform(-0.5, 0.5, 2)
    Ry = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    Rx = np.array([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
    R = Ry @ Rx
    T = rs.uniform(-1, 1, 3)
    fovx = rs.uniform(0.6, 1.4); fovy = rs.uniform(0.5, 1.2)
    primx, primy = (0.5, 0.5) if i < 3 else tuple(rs.uniform(0.4, 0.6, 2))
    wv = torch.tensor(getWorld2View2(R, T, np.array([0.0, 0.0, 0.0]), 1.0)).transpose(0, 1)
    pr = getProjectionMatrix(znear=0.01, zfar=100.0, fovX=fovx, fovY=fovy, primx=primx, primy=primy).transpose(0, 1)
    full = (wv.unsqueeze(0).bmm(pr.unsqueeze(0))).squeeze(0)
    center = wv.inverse()[3, :3]
    cams.append(dict(R=R, T=T, fovx=fovx, fovy=fovy, primx=primx, primy=primy, wv=wv.numpy(), pr=pr.numpy(),
                     full=full.numpy(), center=center.numpy()))
np.savez(os.path.join(HERE, "camera.npz"), **{f"{k}_{i}": np.asarray(c[k]) for i, c in enumerate(cams) for k in c})
print("golden written")

# ---- L1 + SSIM loss (utils/loss_utils.py:17-63, combined as train_post.py:134-140) ----
from utils.loss_utils import l1_loss, ssim          # noqa: E402
gl = torch.Generator().manual_seed(99)
for name, (C_, H_, W_) in {"a": (3, 37, 53), "b": (3, 16, 16), "c": (1, 70, 9)}.items():
    img = torch.rand(C_, H_, W_, generator=gl).requires_grad_(True)
    gt = (img.detach() * 0.7 + 0.3 * torch.rand(C_, H_, W_, generator=gl)).clamp(0, 1)
    lam = 0.2
    Ll1 = l1_loss(img, gt)
    s = ssim(img, gt)
    loss = (1.0 - lam) * Ll1 + lam * (1.0 - s)
    loss.backward()
    np.savez(os.path.join(HERE, f"loss_{name}.npz"), img=img.detach().numpy(), gt=gt.numpy(), lam=lam,
             l1=Ll1.item(), ssim=s.item(), loss=loss.item(), grad=img.grad.numpy())
print("loss golden written")

# ---- sparse Adam (scene/OurAdam.py:249-337 via Adam.step(relevant), as driven by train_single.py:170-178) ----
import importlib.util                                   # noqa: E402
_spec = importlib.util.spec_from_file_location("ref_OurAdam", os.path.join(REF, "scene", "OurAdam.py"))
_mod = importlib.util.module_from_spec(_spec); _spec.loader.exec_module(_mod)
# the reference targets an older torch: this private Optimizer hook was renamed since; it does no arithmetic
if not hasattr(_mod.Adam, "_cuda_graph_capture_health_check"):
    _mod.Adam._cuda_graph_capture_health_check = lambda self: None
ga = torch.Generator().manual_seed(7)
Nrows = 500
params = [torch.randn(Nrows, 3, generator=ga), torch.randn(Nrows, 15, 3, generator=ga), torch.rand(Nrows, 1, generator=ga)]
p0 = [p.clone().numpy() for p in params]
plist = [torch.nn.Parameter(p.clone()) for p in params]
lrs = [1.6e-4, 1.25e-4, 0.05]
opt = _mod.Adam([{"params": [p], "lr": lr} for p, lr in zip(plist, lrs)], lr=0.0, eps=1e-15)
steps = []
for it in range(4):
    rel = torch.nonzero(torch.rand(Nrows, generator=ga) < 0.3).flatten().long()
    grads = [torch.randn(p.shape, generator=ga) * 1e-3 for p in plist]
    for p, g_ in zip(plist, grads):
        p.grad = g_.clone()
    opt.step(rel)
    steps.append(dict(rel=rel.numpy(), grads=[g_.numpy() for g_ in grads], after=[p.detach().clone().numpy() for p in plist]))
out = {"lrs": np.array(lrs), "eps": 1e-15, "n_steps": len(steps)}
for i, p in enumerate(p0):
    out[f"p0_{i}"] = p
for s_i, st in enumerate(steps):
    out[f"rel_{s_i}"] = st["rel"]
    for i in range(3):
        out[f"grad_{s_i}_{i}"] = st["grads"][i]; out[f"after_{s_i}_{i}"] = st["after"][i]
st_ = opt.state[plist[0]]
out["exp_avg_0"] = st_["exp_avg"].numpy(); out["exp_avg_sq_0"] = st_["exp_avg_sq"].numpy()
np.savez(os.path.join(HERE, "sparse_adam.npz"), **out)
print("sparse adam golden written")
