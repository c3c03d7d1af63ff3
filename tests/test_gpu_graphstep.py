"""GPU: the sync-free step (device-side LOD cut, capacity mode, CUDA-graph replay) against the exact
path it mirrors (pipeline.l1_step, fused form) -- same kernels in the same order, so integer outputs
and the image must be bit-identical and the gradients equal up to the order of the atomic sums."""
import numpy as np
import pytest

from h3dgs import synth
from util import rel_err

pytestmark = pytest.mark.gpu


def _scene(skybox=0, leaves=9000, W=480, H=270, seed=3):
    cam = synth.make_camera(W, H)
    lv = synth.cloud_v1(leaves, cam, zmin=2.0, zmax=40.0, seed=seed, scale_k=1.0)
    z = lv["means3D"][:, 2:3]
    lv["scales"] = (4e-3 * np.sqrt(2.0 * z) * np.exp(0.4 * np.random.default_rng(1).standard_normal((z.shape[0], 3)))).astype(np.float32)
    h = synth.build_hierarchy(lv)
    if skybox:
        h = synth.append_skybox(h, skybox)
    return cam, h


def _cams(W, H, n=3):
    rs = np.random.default_rng(2)
    return [synth.make_camera(W, H)] + [synth.yaw_camera(W, H, float(rs.uniform(-15, 15)), rs.uniform(-0.5, 0.5, 3))
                                        for _ in range(n - 1)]


@pytest.mark.parametrize("tau", [0.0, 6.0, 40.0])
def test_device_lod_cut_equals_the_two_call_api(tau):
    import ctypes as C
    import torch
    from gaussian_hierarchy._C import expand_to_size, get_interpolation_weights
    from h3dgs import _lib
    cam, h = _scene()
    N = h["nodes"].shape[0]
    thr = synth.tau_threshold(tau, cam)
    vp = torch.tensor([0.3, -0.2, 1.0], device="cuda")
    nodes, boxes = torch.tensor(h["nodes"], device="cuda"), torch.tensor(h["boxes"], device="cuda")
    z = lambda dt: torch.zeros(N, dtype=dt, device="cuda")
    r, p, nn_, ts, kids = z(torch.int32), z(torch.int32), z(torch.int32), z(torch.float32), z(torch.int32)
    n = expand_to_size(nodes, boxes, thr, vp, torch.zeros(3), r, p, nn_)
    get_interpolation_weights(nn_[:n], thr, nodes, boxes, vp.cpu(), torch.zeros(3), ts, kids)
    r2, p2, nn2, ts2, kids2 = z(torch.int32), z(torch.int32), z(torch.int32), z(torch.float32), z(torch.int32)
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    L = _lib.lib()
    scratch = torch.empty(int(L.h3dgs_expand_scratch_bytes(N)), dtype=torch.uint8, device="cuda")
    thr_dev = torch.full((1,), thr, dtype=torch.float32, device="cuda")
    _lib.check(L.h3dgs_lod_cut(N, nodes.data_ptr(), boxes.data_ptr(), -1.0, thr_dev.data_ptr(), vp.data_ptr(), r2.data_ptr(), p2.data_ptr(),
                               nn2.data_ptr(), ts2.data_ptr(), kids2.data_ptr(), count.data_ptr(), scratch.data_ptr(),
                               torch.cuda.current_stream().cuda_stream))
    assert int(count.item()) == n and n > 0
    for a, b in ((r, r2), (p, p2), (nn_, nn2), (kids, kids2)):
        assert torch.equal(a[:n], b[:n])
    assert torch.equal(ts[:n].view(torch.int32), ts2[:n].view(torch.int32))
    assert bool((r2[n:] == -1).all())


def _exact_step(scene, dcam, bg, gt, thr):
    import torch
    from h3dgs import pipeline
    loss, radii, n = pipeline.l1_step(scene, dcam, bg, gt, thr, fused=True)
    grads = {k: getattr(scene, k).grad.clone() for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    with torch.no_grad():
        img = pipeline.render_hier_fused(scene, dcam, bg, thr)[0]
    from diff_gaussian_rasterization import _C
    return float(loss.item()), radii.clone(), n, grads, img.clone(), _C.last_num_rendered()


@pytest.mark.parametrize("skybox", [0, 200])
@pytest.mark.parametrize("capture", [False, True])
def test_sync_free_step_equals_exact_step(skybox, capture):
    import torch
    from h3dgs import pipeline
    from h3dgs.graphstep import GraphedStep
    cam, h = _scene(skybox=skybox)
    cams = _cams(cam.W, cam.H)
    thr = synth.tau_threshold(6.0, cam)
    scene = pipeline.Scene(h)
    bg = torch.tensor([0.2, 0.1, 0.3], device="cuda")
    g = torch.Generator(device="cpu").manual_seed(5)
    gts = [torch.rand((3, cam.H, cam.W), generator=g).cuda() for _ in cams]
    dcams = [pipeline.DeviceCamera(c) for c in cams]
    gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, bin_capacity=1 << 20, sort_capacity=4096,
                     capture=False)
    gs.set_camera(dcams[0]); gs.gt.copy_(gts[0])
    if capture:
        gs.capture()
        assert gs.launches_per_step >= 10
    for v in (0, 1, 2, 1):
        loss, radii, n, grads, img, D = _exact_step(scene, dcams[v], bg, gts[v], thr)
        gs.step(dcams[v], gts[v])
        st = gs.status()
        assert not st["overflow"]
        assert st["rows"] == n + skybox and st["D"] == D and 0 < st["longest_list"] <= 4096
        assert abs(st["loss"] - loss) < 1e-7
        assert torch.equal(gs.image, img)
        P = n + skybox
        assert torch.equal(gs.radii[:P], radii) and bool((gs.radii[P:] == 0).all())
        for k, ref in grads.items():
            e = rel_err(gs.grads[k].cpu().numpy(), ref.cpu().numpy())
            assert e < 5e-6, (v, k, e)               # same partial sums, different atomic order (2.1e-6 seen on a B200)
    # a new LOD threshold between replays (train_post.py:66-74 draws one per step): it lives on the device
    thr2 = synth.tau_threshold(15.0, cam)
    loss, radii, n, grads, img, D = _exact_step(scene, dcams[0], bg, gts[0], thr2)
    gs.set_threshold(thr2)
    gs.step(dcams[0], gts[0])
    st = gs.status()
    assert not st["overflow"] and st["rows"] == n + skybox and st["D"] == D and torch.equal(gs.image, img)


def test_capacity_overflow_is_flagged_and_harmless():
    import torch
    from h3dgs import pipeline
    from h3dgs.graphstep import GraphedStep
    cam, h = _scene()
    thr = synth.tau_threshold(6.0, cam)
    scene = pipeline.Scene(h)
    bg = torch.tensor([0.2, 0.1, 0.3], device="cuda")
    dcam = pipeline.DeviceCamera(cam)
    gt = torch.rand((3, cam.H, cam.W), device="cuda")
    loss, radii, n, grads, img, D = _exact_step(scene, dcam, bg, gt, thr)
    for kw in (dict(bin_capacity=D // 2, sort_capacity=4096), dict(bin_capacity=1 << 20, sort_capacity=32),
               dict(bin_capacity=1 << 20, sort_capacity=4096, row_capacity=n // 2)):
        gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, capture=False, **kw)
        gs.step(dcam, gt)
        st = gs.status()
        assert st["overflow"] and st["D"] == D if "row_capacity" not in kw else st["overflow"]
        if "row_capacity" not in kw:
            # the frame was turned into an empty one: background only, zero gradients
            assert torch.equal(gs.image, bg.view(3, 1, 1).expand_as(gs.image))
            assert all(float(v.abs().sum()) == 0.0 for v in gs.grads.values())
    # exactly fitting capacities are not an overflow
    gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, capture=False, bin_capacity=D,
                     sort_capacity=4096, row_capacity=n)
    gs.step(dcam, gt)
    assert not gs.status()["overflow"] and torch.equal(gs.image, img)


def test_capacity_mode_rejects_debug_and_bad_capacities():
    import ctypes as C
    import torch
    from h3dgs import _lib, pipeline
    from h3dgs.graphstep import GraphedStep
    cam, h = _scene(leaves=500)
    scene = pipeline.Scene(h)
    gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, torch.zeros(3, device="cuda"),
                     synth.tau_threshold(6.0, cam), capture=False)
    gs.args.sort_capacity = 8193
    with pytest.raises(RuntimeError, match="capacit"):
        gs.step(pipeline.DeviceCamera(cam), torch.zeros((3, cam.H, cam.W), device="cuda"))
    gs.args.sort_capacity, gs.args.debug = 4096, 1
    with pytest.raises(RuntimeError, match="debug"):
        gs.step(pipeline.DeviceCamera(cam), torch.zeros((3, cam.H, cam.W), device="cuda"))
