"""CPU: the product's Python layer end to end -- drop-in packages, autograd Function, h3dgs.pipeline, the
sync-free step of h3dgs.graphstep (without graph capture) and the tile-sharded mode over gloo -- on CPU tensors,
with the emulation build of the kernels standing in for libh3dgs.so (tests/emul/fake_device.py).  Everything
except CUDA graph capture and NCCL itself is exercised before any GPU time is spent."""
import os
import sys

import numpy as np
import pytest
import torch

from h3dgs import synth
from util import rel_err

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emul"))


@pytest.fixture(scope="module")
def emu_so(tmp_path_factory):
    from build_emu import build
    return build(str(tmp_path_factory.mktemp("h3dgs_emu_glue")))


def _scene(skybox=0, leaves=2500, W=160, H=112):
    cam = synth.make_camera(W, H)
    lv = synth.cloud_v1(leaves, cam, zmin=2.0, zmax=40.0, seed=3, scale_k=1.0)
    z = lv["means3D"][:, 2:3]
    lv["scales"] = (1.2e-2 * np.sqrt(2.0 * z) * np.exp(0.4 * np.random.default_rng(1).standard_normal((z.shape[0], 3)))).astype(np.float32)
    h = synth.build_hierarchy(lv)
    if skybox:
        h = synth.append_skybox(h, skybox)
    return cam, h


def _cams(W, H, n=3):
    rs = np.random.default_rng(2)
    return [synth.make_camera(W, H)] + [synth.yaw_camera(W, H, float(rs.uniform(-15, 15)), rs.uniform(-0.5, 0.5, 3))
                                        for _ in range(n - 1)]


@pytest.mark.parametrize("skybox", [0, 60])
def test_public_api_step_and_sync_free_step(emu_so, skybox):
    from fake_device import cpu_as_device
    from test_gpu_pipeline import _oracle_hier_step
    cam, h = _scene(skybox=skybox)
    thr = synth.tau_threshold(6.0, cam)
    gt = np.random.default_rng(2).uniform(0, 1, (3, cam.H, cam.W)).astype(np.float32)
    n_ref, f, gref = _oracle_hier_step(h, cam, thr, gt)
    with cpu_as_device(emu_so):
        from h3dgs import pipeline
        from h3dgs.graphstep import GraphedStep
        scene = pipeline.Scene(h, device="cpu")
        dcam = pipeline.DeviceCamera(cam, device="cpu")
        bg, gtd = torch.zeros(3), torch.tensor(gt)
        # the reference-facing path: expand_to_size -> weights -> (gather/lerp) -> GaussianRasterizer -> L1 -> backward
        for fused in (False, True):
            loss, radii, n = pipeline.l1_step(scene, dcam, bg, gtd, thr, fused=fused)
            assert n == n_ref and np.array_equal(radii.numpy(), f["radii"])
            assert abs(loss.item() - np.abs(f["color"] - gt).mean()) < 1e-6
            for name in ("means3D", "scales", "shs", "opacities", "rotations"):
                assert rel_err(getattr(scene, name).grad.numpy(), gref[name]) < 5e-5, (fused, name)
        exact = {k: getattr(scene, k).grad.clone() for k in ("means3D", "scales", "shs", "opacities", "rotations")}
        with torch.no_grad():
            img = pipeline.render_hier_fused(scene, dcam, bg, thr)[0].clone()
        from diff_gaussian_rasterization import _C as rc
        D = rc.last_num_rendered()
        # the sync-free step (device-side cut, capacity mode); graph capture itself needs a GPU
        gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, bin_capacity=4 * D + 1000, sort_capacity=4096,
                         capture=False)
        for _ in range(2):
            gs.step(dcam, gtd)
            st = gs.status()
            assert not st["overflow"] and st["rows"] == n_ref + skybox and st["D"] == D and st["longest_list"] > 0
            assert abs(st["loss"] - loss.item()) < 1e-7
            assert torch.equal(gs.image, img)
            assert torch.equal(gs.radii[:n_ref + skybox], radii) and bool((gs.radii[n_ref + skybox:] == 0).all())
            for k, ref in exact.items():
                assert rel_err(gs.grads[k].numpy(), ref.numpy()) < 2e-6, k
        # targets uploaded on another stream alternate between the two target buffers; a step reads the buffer its upload wrote
        gt_b = torch.tensor(np.random.default_rng(9).uniform(0, 1, (3, cam.H, cam.W)).astype(np.float32))
        loss_b = pipeline.l1_step(scene, dcam, bg, gt_b, thr, fused=True)[0].item()
        side = torch.cuda.Stream() if hasattr(torch.cuda, "Stream") else None
        for target, want in ((gtd, loss.item()), (gt_b, loss_b), (gt_b, loss_b), (gtd, loss.item())):
            slot_before = gs._upload_slot
            ready = gs.upload_target(target, side)
            assert gs._upload_slot == 1 - slot_before and gs._pending[0] == slot_before
            gs.set_camera(dcam)
            gs.step(gt_ready=ready)
            assert gs._pending is None and abs(gs.status()["loss"] - want) < 1e-7
        gs.step(dcam, gt_b)                                  # step(gt=) goes through buffer 0 whatever was uploaded before
        assert abs(gs.status()["loss"] - loss_b) < 1e-7 and torch.equal(gs.gt, gt_b)
        # a new threshold and a new camera between steps
        thr2 = synth.tau_threshold(15.0, cam)
        cam2 = pipeline.DeviceCamera(_cams(cam.W, cam.H)[1], device="cpu")
        loss2, radii2, n2 = pipeline.l1_step(scene, cam2, bg, gtd, thr2, fused=True)
        gs.set_threshold(thr2)
        gs.step(cam2, gtd)
        st = gs.status()
        assert not st["overflow"] and st["rows"] == n2 + skybox and abs(st["loss"] - loss2.item()) < 1e-7
        # capacities that do not fit are reported, not crashed on
        small = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, bin_capacity=D // 3, sort_capacity=4096,
                            capture=False)
        small.step(dcam, gtd)
        assert small.status()["overflow"]
        rows = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, bin_capacity=D + 100, sort_capacity=4096,
                           row_capacity=(n_ref + skybox) // 2, capture=False)
        rows.step(dcam, gtd)
        assert rows.status()["overflow"]


def _sharded_worker(rank, world, port, so, out):
    sys.path.insert(0, os.path.join(HERE, "emul")); sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "hierarchical-3d-gaussians_b200"))
    import torch.distributed as dist
    from fake_device import cpu_as_device
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cam, h = _scene()
    thr = synth.tau_threshold(6.0, cam)
    gt = torch.tensor(np.random.default_rng(2).uniform(0, 1, (3, cam.H, cam.W)).astype(np.float32))
    ok, errs = True, []
    with cpu_as_device(so):
        from h3dgs import pipeline, dist as hd
        from h3dgs.graphstep import GraphedStep
        scene = pipeline.Scene(h, device="cpu")
        dcam = pipeline.DeviceCamera(cam, device="cpu")
        bg = torch.tensor([0.1, 0.2, 0.3])
        loss1, radii1, n1 = pipeline.l1_step(scene, dcam, bg, gt, thr)
        g1 = {k: getattr(scene, k).grad.clone() for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        with torch.no_grad():
            img1 = pipeline.render_hier_fused(scene, dcam, bg, thr)[0].clone()
        # tile-sharded step of h3dgs.dist (autograd Function, all-gather of slabs, reduction of the [P,10] sums)
        sh = hd.TileSharder(world, rank, "cpu")
        loss2, radii2, n2 = sh.l1_step(scene, dcam, bg, gt, thr)
        ok = ok and n1 == n2 and torch.equal(radii1, radii2) and abs(loss1.item() - loss2.item()) < 1e-7
        for k, ref in g1.items():
            t_ = getattr(scene, k).grad.clone()
            dist.all_reduce(t_)                               # gradients come back sharded by rendered row
            errs.append(float((ref - t_).abs().max() / ref.abs().max().clamp_min(1e-30)))
        # the sync-free sharded step (no capture)
        gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, bin_capacity=1 << 18, sort_capacity=4096,
                         world=world, rank=rank, capture=False)
        gs.step(dcam, gt)
        st = gs.status()
        ok = ok and not st["overflow"] and st["rows"] == n1 and abs(st["loss"] - loss1.item()) < 1e-7
        ok = ok and torch.equal(gs.image, img1)
        for k, ref in g1.items():
            t_ = gs.grads[k].clone()
            dist.all_reduce(t_)
            errs.append(float((ref - t_).abs().max() / ref.abs().max().clamp_min(1e-30)))
    ok = ok and max(errs) < 1e-5
    res = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        torch.save((bool(res.item() == 1.0), errs), out)
    dist.barrier(); dist.destroy_process_group()


def test_tile_sharded_modes_over_gloo(emu_so, tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "r.pt")
    mp.spawn(_sharded_worker, args=(2, port, emu_so, out), nprocs=2, join=True)
    ok, errs = torch.load(out)
    assert ok, errs
