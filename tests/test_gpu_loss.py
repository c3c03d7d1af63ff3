"""GPU: fused L1 + SSIM loss -- PINNED to the reference's own code: golden loss values and
gradients were produced by importing /root/reference/utils/loss_utils.py (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_matches_reference_loss_utils(golden_dir, name):
    import torch
    from h3dgs.loss import fused_l1_ssim
    z = np.load(os.path.join(golden_dir, f"loss_{name}.npz"))
    img = torch.tensor(z["img"], device="cuda", requires_grad=True)
    gt = torch.tensor(z["gt"], device="cuda")
    loss, l1, ssim = fused_l1_ssim(img, gt, float(z["lam"]))
    assert abs(l1.item() - float(z["l1"])) < 1e-6 and abs(ssim.item() - float(z["ssim"])) < 2e-6
    assert abs(loss.item() - float(z["loss"])) < 1e-6
    loss.backward()
    g = img.grad.cpu().numpy()
    assert np.abs(g - z["grad"]).max() < 1e-5 * np.abs(z["grad"]).max()


def _torch_reference(img, gt, lam):
    """the reference formulation (grouped conv2d, 121 taps) restated for on-GPU comparison at 1080p"""
    import math
    import torch
    import torch.nn.functional as F
    g = torch.tensor([math.exp(-(x - 5) ** 2 / (2 * 1.5 ** 2)) for x in range(11)])
    g = (g / g.sum()).unsqueeze(1)
    C = img.shape[0]
    w = g.mm(g.t()).float()[None, None].expand(C, 1, 11, 11).contiguous().to(img.device)
    a, b = img[None], gt[None]
    mu1, mu2 = F.conv2d(a, w, padding=5, groups=C), F.conv2d(b, w, padding=5, groups=C)
    s1 = F.conv2d(a * a, w, padding=5, groups=C) - mu1 * mu1
    s2 = F.conv2d(b * b, w, padding=5, groups=C) - mu2 * mu2
    s12 = F.conv2d(a * b, w, padding=5, groups=C) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()
    return (1 - lam) * (img - gt).abs().mean() + lam * (1 - ssim)


def test_full_hd_against_conv2d_formulation_and_timing():
    import torch
    from h3dgs.loss import fused_l1_ssim
    gen = torch.Generator(device="cuda").manual_seed(0)
    gt = torch.rand((3, 1080, 1920), device="cuda", generator=gen)
    base = (gt * 0.8 + 0.2 * torch.rand((3, 1080, 1920), device="cuda", generator=gen)).clamp(0, 1)
    img1 = base.clone().requires_grad_(True)
    img2 = base.clone().requires_grad_(True)
    loss = fused_l1_ssim(img1, gt, 0.2)[0]
    loss.backward()
    ref = _torch_reference(img2, gt, 0.2)
    ref.backward()
    assert abs(loss.item() - ref.item()) < 2e-6
    assert float((img1.grad - img2.grad).abs().max() / img2.grad.abs().max()) < 2e-5

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    def ours():
        img1.grad = None
        fused_l1_ssim(img1, gt, 0.2)[0].backward()

    def theirs():
        img2.grad = None
        _torch_reference(img2, gt, 0.2).backward()
    t_ours, t_ref = timeit(ours), timeit(theirs)
    print(f"\nL1+SSIM fwd+bwd @1080p: fused {t_ours:.3f} ms vs conv2d formulation {t_ref:.3f} ms")
    assert t_ours < t_ref
