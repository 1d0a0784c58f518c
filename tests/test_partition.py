"""N > 1 host logic on CPU: tile partition + the framebuffer reduce over gloo, world size 2.

The per-rank images come from the oracle's partitioned render (the same tile rule as igb200_set_partition), so the test
checks what bench.py relies on at N > 1: disjoint supports, and reduce(sum) == the single-device image exactly.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from ignis_b200.partition import TileGather, reduce_framebuffer, tile_owner
    from ignis_b200.scene import load_scene
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h, spi = 96, 80, 2
    t = load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json"))
    o = Oracle(t)
    fb = np.zeros((h, w, 3), np.float32)
    for it in range(2):
        o.render(w, h, spi=spi, iteration=it, fb=fb, threads=2, partition=(rank, world, 32))
    own = tile_owner(w, h, world, 32) == rank
    assert not fb[~own].any(), "a rank wrote outside its tiles"
    # the two forms of the exchange: gather of the rank's own tiles, sum of whole frames
    g = TileGather(w, h, rank, world, 32)
    full = g.run(torch.from_numpy(fb.copy()).reshape(-1), out=torch.zeros(h * w * 3))
    ft = torch.from_numpy(fb)
    reduce_framebuffer(ft, dst=0)
    if rank == 0:
        assert torch.equal(full.reshape(h, w, 3), ft), "gather of tiles != sum of frames"
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), ft.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_tile_owner_covers_frame_once():
    from ignis_b200.partition import local_ray_domain, tile_owner
    for (w, h, world) in [(1920, 1080, 8), (100, 70, 3), (31, 33, 2), (64, 64, 1)]:
        own = tile_owner(w, h, world)
        assert own.shape == (h, w) and own.min() >= 0 and own.max() < world
        counts = [(own == r).sum() for r in range(world)]
        assert sum(counts) == w * h
        for r in range(world):
            assert local_ray_domain(w, h, 4, r, world) >= counts[r] * 4
    own8 = tile_owner(1920, 1080, 8)
    frac = np.bincount(own8.ravel(), minlength=8) / own8.size
    assert frac.max() / frac.min() < 1.05   # round-robin 32x32 tiles balance the pixel count within 5 %


@pytest.mark.timeout(300)
def test_gloo_world2_reduce_equals_single_device(tmp_path):
    import torch.multiprocessing as mp
    from ignis_b200.scene import load_scene
    from oracle.oracle import Oracle
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "reduced.npy")
    w, h, spi = 96, 80, 2
    o = Oracle(load_scene(os.path.join(ROOT, "scenes", "diamond_scene.json")))
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(2):
        o.render(w, h, spi=spi, iteration=it, fb=ref, threads=2)
    assert ref.sum() > 0
    np.testing.assert_array_equal(got, ref)


def test_tile_gather_single_process_and_index_sets():
    """World 1 needs no process group: the frame comes back unchanged. The index sets of all ranks partition the frame."""
    import torch
    from ignis_b200.partition import TileGather, tile_owner
    w, h = 100, 70
    fb = torch.arange(w * h * 3, dtype=torch.float32)
    g = TileGather(w, h, 0, 1, 32)
    assert torch.equal(g.run(fb.clone(), out=torch.zeros_like(fb)), fb)
    world = 3
    own = tile_owner(w, h, world, 32).ravel()
    seen = np.zeros(w * h, np.int32)
    for r in range(world):
        tg = TileGather.__new__(TileGather)   # only the bookkeeping: no collective is run here
        TileGather.__init__(tg, w, h, r, world, 32, dst=r)
        idx = tg.idx[r].numpy()
        assert len(idx) == tg.counts[r] <= tg.n_max and (own[idx] == r).all()
        seen[idx] += 1
    assert (seen == 1).all()
