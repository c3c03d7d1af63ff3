"""GPU (needs >= 2 devices; skipped otherwise): the tile-sharded step over NCCL reproduces the
single-GPU step -- image bit-identical (same per-tile order), gradients to fp32 sum order."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _spawn(fn, args, world, timeout_s=240):
    """mp.spawn with a deadline: a multi-rank hang (a peer that never arrives, a stuck collective) must fail this test
    instead of stalling the whole GPU tier; the children are killed by their exact pids."""
    import time
    import torch.multiprocessing as mp
    ctx = mp.spawn(fn, args=args, nprocs=world, join=False)
    deadline = time.time() + timeout_s
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"{fn.__name__}: {world} ranks did not finish within {timeout_s} s")


def _worker(rank, world, port, out):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "hierarchical-3d-gaussians_b200")); sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    from h3dgs import synth, pipeline, dist as hd
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    cam = synth.make_camera(640, 360)
    leaves = synth.cloud_v1(20000, cam, zmin=2.0, zmax=40.0, seed=3, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (5e-3 * np.sqrt(2.0 * z) * np.ones((1, 3))).astype(np.float32)
    h = synth.build_hierarchy(leaves)
    thr = synth.tau_threshold(6.0, cam)
    scene = pipeline.Scene(h, device=dev)
    dcam = pipeline.DeviceCamera(cam, device=dev)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    gt = torch.rand((3, cam.H, cam.W), generator=torch.Generator().manual_seed(1)).to(dev)
    loss1, radii1, n1 = pipeline.l1_step(scene, dcam, bg, gt, thr)
    g1 = [p.grad.clone() for p in scene.params()]
    with torch.no_grad():
        img1 = pipeline.render_hier_fused(scene, dcam, bg, thr)[0]
    sh = hd.TileSharder(world, rank, dev)
    loss2, radii2, n2 = sh.l1_step(scene, dcam, bg, gt, thr)
    g2 = [p.grad.clone() for p in scene.params()]
    for t_ in g2:                              # gradients come back sharded by rendered row: their sum is the full gradient
        dist.all_reduce(t_, op=dist.ReduceOp.SUM)
    with torch.no_grad():
        img2 = sh.render(scene, dcam, bg, thr)[0]
    ok = n1 == n2 and torch.equal(radii1, radii2) and torch.equal(img1, img2) and abs(loss1.item() - loss2.item()) < 1e-7
    errs = [float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)) for a, b in zip(g1, g2)]
    ok = ok and max(errs) < 1e-5
    res = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        torch.save((bool(res.item() == 1.0), errs), out)
    dist.barrier(); dist.destroy_process_group()


def _graph_worker(rank, world, port, out, peer=False):
    """the same comparison for the sync-free step replayed from CUDA graphs (collectives captured); peer=True: the
    collectives fused into the blend kernels over peer memory (forward stores into every rank's image, backward adds
    into the owner's accumulator, device-side barriers)"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "hierarchical-3d-gaussians_b200")); sys.path.insert(0, os.path.join(root, "tests"))
    import torch
    import torch.distributed as dist
    from h3dgs import synth, pipeline
    from h3dgs.graphstep import GraphedStep
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    cam = synth.make_camera(640, 360)
    leaves = synth.cloud_v1(20000, cam, zmin=2.0, zmax=40.0, seed=3, scale_k=1.0)
    z = leaves["means3D"][:, 2:3]
    leaves["scales"] = (5e-3 * np.sqrt(2.0 * z) * np.ones((1, 3))).astype(np.float32)
    h = synth.build_hierarchy(leaves)
    thr = synth.tau_threshold(6.0, cam)
    scene = pipeline.Scene(h, device=dev)
    rs = np.random.default_rng(2)
    cams = [cam] + [synth.yaw_camera(cam.W, cam.H, float(rs.uniform(-15, 15)), rs.uniform(-0.5, 0.5, 3)) for _ in range(2)]
    dcams = [pipeline.DeviceCamera(c, device=dev) for c in cams]
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    gen = torch.Generator().manual_seed(1)
    gts = [torch.rand((3, cam.H, cam.W), generator=gen).to(dev) for _ in cams]
    gs = GraphedStep(scene, cam.W, cam.H, cam.tanfovx, cam.tanfovy, bg, thr, bin_capacity=1 << 20, sort_capacity=4096,
                     world=world, rank=rank, capture=False, peer=peer, cyclic_log2=7)
    gs.set_camera(dcams[0]); gs.gt.copy_(gts[0])
    gs.capture()
    ok, errs = True, []
    for v in (0, 1, 2, 0):
        loss1, radii1, n1 = pipeline.l1_step(scene, dcams[v], bg, gts[v], thr)
        g1 = {k: getattr(scene, k).grad.clone() for k in ("means3D", "shs", "opacities", "scales", "rotations")}
        with torch.no_grad():
            img1 = pipeline.render_hier_fused(scene, dcams[v], bg, thr)[0]
        gs.step(dcams[v], gts[v])
        st = gs.status()
        ok = ok and not st["overflow"] and st["rows"] == n1 and abs(st["loss"] - loss1.item()) < 1e-7
        ok = ok and torch.equal(gs.image, img1) and torch.equal(gs.radii[:n1], radii1)
        for k, ref in g1.items():
            t_ = gs.grads[k].clone()
            dist.all_reduce(t_, op=dist.ReduceOp.SUM)          # sharded by rendered row: the sum is the full gradient
            errs.append(float((ref - t_).abs().max() / ref.abs().max().clamp_min(1e-30)))
    ok = ok and max(errs) < 1e-5
    if peer:
        ok = ok and not gs.arena.timed_out()
    res = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        torch.save((bool(res.item() == 1.0), errs), out)
    if peer:
        gs.arena.close()
    dist.barrier(); dist.destroy_process_group()


def test_graphed_peer_step_equals_single_gpu(tmp_path):
    """peer mode: no NCCL inside the step"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 4 if torch.cuda.device_count() >= 4 else 2
    out = str(tmp_path / "r.pt")
    _spawn(_graph_worker, (world, _free_port(), out, True), world)
    ok, errs = torch.load(out)
    assert ok, errs


def test_graphed_sharded_step_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    out = str(tmp_path / "r.pt")
    _spawn(_graph_worker, (world, _free_port(), out), world)
    ok, errs = torch.load(out)
    assert ok, errs


def test_tile_sharded_step_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    out = str(tmp_path / "r.pt")
    _spawn(_worker, (world, _free_port(), out), world)
    ok, errs = torch.load(out)
    assert ok, errs
