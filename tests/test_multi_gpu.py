"""Multi-GPU through the C ABI alone (include/igb200.h igb200_comm_*): one process per GPU, NCCL inside the device library, no
torch.distributed anywhere. Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on a single-GPU box. tests/test_partition.py covers the host
logic of the same partition at world size 2 on CPU (gloo)."""
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _rank_stream(rank, world, conn, scene, w, h, spi, iters, fuse):
    """Frame streaming across ranks: every iteration's gathered frame arrives on rank 0, in order, while later iterations render."""
    sys.path.insert(0, ROOT)
    from ignis_b200.device import Runtime
    from ignis_b200.scene import load_scene
    try:
        uid = conn.recv()
        t = load_scene(os.path.join(ROOT, "scenes", scene))
        with Runtime(t, w, h, spi=spi, cuda_device=rank) as rt:
            rt.device.commInit(rank, world, uid, 32)
            share = 0
            if isinstance(fuse, tuple):
                fuse, share = fuse
            rt.device.setOption("fuse", fuse)
            if share:
                rt.device.frameStreamShare(share)   # frames in shared host memory: every rank writes its own tiles, no exchange between the GPUs
            rt.device.frameStreamBegin(32)
            out = {}
            for it in range(iters):
                rt.step()
                while (f := rt.device.frameStreamNext(0)) is not None:
                    out[f[0]] = f[1].copy()
            while (f := rt.device.frameStreamNext(2)) is not None:
                out[f[0]] = f[1].copy()
            st = rt.device.getStatistics()
            rt.device.frameStreamEnd()
        conn.send(("ok", [out[k] for k in sorted(out)] if rank == 0 else [], (st["CameraRayCount"], st["ShadowRayCount"], st["BounceRayCount"])))
    except Exception as e:   # noqa: BLE001
        conn.send(("error", repr(e), None))


def _rank_main(rank, world, conn, scene, w, h, spi, iters, gather_without_sync):
    sys.path.insert(0, ROOT)
    from ignis_b200.device import B200Device, Runtime
    from ignis_b200.scene import load_scene
    try:
        uid = conn.recv()
        t = load_scene(os.path.join(ROOT, "scenes", scene))
        with Runtime(t, w, h, spi=spi, cuda_device=rank) as rt:
            rt.device.commInit(rank, world, uid, 32)
            out = []
            for it in range(iters):
                rt.step()
                if gather_without_sync or it == iters - 1:
                    # no igb200_sync in between: the gather itself is ordered behind the deferred tail of the render
                    _, host = rt.device.commGatherFramebuffer("", to_host=True)
                    if rank == 0:
                        out.append(host.copy())
            st = rt.device.getStatistics()
        conn.send(("ok", out, (st["CameraRayCount"], st["ShadowRayCount"], st["BounceRayCount"])))
    except Exception as e:   # noqa: BLE001
        conn.send(("error", repr(e), None))


def _run(world, scene, w, h, spi, iters, gather_without_sync=True, target=None, extra=None):
    from ignis_b200.device import B200Device
    ctx = mp.get_context("spawn")
    pipes, procs = [], []
    for r in range(world):
        a, b = ctx.Pipe()
        p = ctx.Process(target=target or _rank_main, args=(r, world, b, scene, w, h, spi, iters, gather_without_sync if extra is None else extra))
        p.start()
        pipes.append(a)
        procs.append(p)
    uid = B200Device.commUniqueId()
    for a in pipes:
        a.send(uid)
    res = [a.recv() if a.poll(600) else ("error", "timeout", None) for a in pipes]
    for p in procs:
        p.join(30)
        if p.is_alive():
            p.kill()
    for r in res:
        assert r[0] == "ok", r[1]
    return res


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("scene,w,h,spi", [("diamond_scene.json", 480, 270, 4), ("many_point_lights.json", 320, 320, 1)])
def test_gathered_frame_equals_oracle(world, scene, w, h, spi):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    from ignis_b200.scene import load_scene
    from oracle.oracle import Oracle
    iters = 3
    res = _run(world, scene, w, h, spi, iters)
    o = Oracle(load_scene(os.path.join(ROOT, "scenes", scene)))
    ref = np.zeros((h, w, 3), np.float32)
    frames = res[0][1]
    assert len(frames) == iters
    for it in range(iters):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
        err = float(np.linalg.norm((frames[it] - ref).ravel()) / np.linalg.norm(ref.ravel()))
        assert err <= 1e-4, (it, err)            # every gathered frame, taken right behind an asynchronous render
    counts = np.sum([r[2] for r in res], axis=0)
    assert tuple(int(x) for x in counts) == tuple(int(x) for x in o.counters)


@pytest.mark.parametrize("world,fuse", [(2, 1), (2, 4), (8, 8), (2, (1, 0x1b200)), (2, (4, 0x1b201)), (8, (8, 0x1b202))])
def test_streamed_gathered_frames_equal_oracle(world, fuse):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    if isinstance(fuse, tuple):
        fuse = (fuse[0], fuse[1] + (os.getpid() & 0xFFF) * 16)   # a key of this test run
    from ignis_b200.scene import load_scene
    from oracle.oracle import Oracle
    scene, w, h, spi, iters = "diamond_scene.json", 480, 270, 4, 18
    res = _run(world, scene, w, h, spi, iters, target=_rank_stream, extra=fuse)
    frames = res[0][1]
    assert len(frames) == iters
    o = Oracle(load_scene(os.path.join(ROOT, "scenes", scene)))
    ref = np.zeros((h, w, 3), np.float32)
    for it in range(iters):
        o.render(w, h, spi=spi, iteration=it, fb=ref)
        err = float(np.linalg.norm((frames[it] - ref).ravel()) / np.linalg.norm(ref.ravel()))
        assert err <= 1e-4, (it, err)
    counts = np.sum([r[2] for r in res], axis=0)
    assert tuple(int(x) for x in counts) == tuple(int(x) for x in o.counters)
