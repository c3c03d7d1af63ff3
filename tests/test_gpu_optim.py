"""GPU: sparse Adam PINNED to the reference's scene/OurAdam.py: four steps with random
`relevant` row sets replayed from golden vectors produced by that file (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sparse_adam_matches_reference_ouradam(golden_dir):
    import torch
    from h3dgs.optim import Adam
    z = np.load(os.path.join(golden_dir, "sparse_adam.npz"))
    params = [torch.nn.Parameter(torch.tensor(z[f"p0_{i}"], device="cuda")) for i in range(3)]
    opt = Adam([{"params": [p], "lr": float(lr)} for p, lr in zip(params, z["lrs"])], lr=0.0, eps=float(z["eps"]))
    mine_prev = [z[f"p0_{i}"] for i in range(3)]
    for s in range(int(z["n_steps"])):
        for i, p in enumerate(params):
            p.grad = torch.tensor(z[f"grad_{s}_{i}"], device="cuda")
        opt.step(torch.tensor(z[f"rel_{s}"], device="cuda"))
        for i, p in enumerate(params):
            ref = z[f"after_{s}_{i}"]
            got = p.detach().cpu().numpy()
            # same fp32 operation sequence as the reference's elementwise kernels: agree to the last few ulp
            assert np.abs(got - ref).max() <= 4e-7 * max(1.0, np.abs(ref).max()), (s, i, np.abs(got - ref).max())
            rows = np.setdiff1d(np.arange(ref.shape[0]), z[f"rel_{s}"])
            assert np.array_equal(got[rows], mine_prev[i][rows])   # rows outside `relevant` are not touched at all
            mine_prev[i] = got
    st = opt.state[params[0]]
    assert np.abs(st["exp_avg"].cpu().numpy() - z["exp_avg_0"]).max() <= 1e-9
    assert np.abs(st["exp_avg_sq"].cpu().numpy() - z["exp_avg_sq_0"]).max() <= 1e-12
    assert float(st["step"]) == float(z["n_steps"])
