"""CPU, world_size 2, gloo: the host logic of the tile-sharded mode (h3dgs/dist.py) --
slab packing, the single all-gather + unpack, and the [P,10] gradient-sum reduction --
reproduces the single-process result exactly."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, H, W, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "hierarchical-3d-gaussians_b200"))
    from h3dgs import dist as hd
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    full = torch.rand((3, H, W), generator=g)
    # each rank "renders" only its tile rows, in the kernel's packed layout
    packed = hd.pack_rows(full, world, rank)
    assert packed.shape[0] == hd.owned_rows(H, world, rank)
    img = hd.gather_image(packed, H, W, world)
    ok_img = torch.equal(img, full)
    # gradient sums: each rank holds a partial [P,10]; padded scratch like the C-ABI's
    P = 1001                                  # not a multiple of the world size
    parts = torch.rand((world, P, 10), generator=g)
    scratch = hd.accum_scratch(P, world, "cpu")
    scratch.view(torch.float32)[: P * 10] = parts[rank].flatten()
    hd.reduce_accum(scratch, P, world, rank)
    lo, hi = hd.row_block(P, world, rank)
    ok_acc = torch.allclose(scratch.view(torch.float32)[lo * 10: hi * 10].view(hi - lo, 10), parts.sum(0)[lo:hi],
                            rtol=1e-6, atol=1e-7)
    ok_acc = ok_acc and sum(b[1] - b[0] for b in (hd.row_block(P, world, r) for r in range(world))) == P
    if rank == 0:
        torch.save((ok_img, ok_acc), out)
    dist.barrier(); dist.destroy_process_group()
    assert ok_img and ok_acc


@pytest.mark.parametrize("H,W", [(1080, 64), (100, 48), (16, 32)])
def test_tile_shard_exchange_world2(tmp_path, H, W):
    out = str(tmp_path / "r.pt")
    mp.spawn(_worker, args=(2, _free_port(), H, W, out), nprocs=2, join=True)
    assert torch.load(out) == (True, True)


def test_row_ownership_partitions_all_tile_rows():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hierarchical-3d-gaussians_b200"))
    from h3dgs import dist as hd
    for H in (16, 17, 1080, 2160, 100):
        gy = (H + 15) // 16
        for world in (1, 2, 3, 4, 8):
            assert sum(hd.owned_rows(H, world, r) for r in range(world)) == gy
            assert max(hd.owned_rows(H, world, r) for r in range(world)) == hd.rows_per_rank(H, world)
