"""BVH8 construction and the on-disk BVH cache (reference: src/runtime/shape/TriMeshProvider.cpp:255-351,
src/runtime/bvh/NArityBvh.h:94-143). The host builder and the cache need no GPU; the GPU builder is checked against the same
structural validator (igb200_test_bvh_build) and against the host builder's tree quality."""
import numpy as np
import pytest

from ignis_b200 import device


def random_boxes(n, seed, clustered=False):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-10, 10, (n, 3)).astype(np.float32)
    if clustered:                       # many primitives share a Morton cell: duplicate keys, deep runs
        c = (c // 5 * 5).astype(np.float32) + rng.normal(0, 1e-4, (n, 3)).astype(np.float32)
    e = rng.uniform(0, 0.3, (n, 3)).astype(np.float32)
    return np.concatenate([c - e, c + e], axis=1)


@pytest.mark.parametrize("n,clustered", [(1, False), (4, False), (5, False), (33, False), (1000, False), (20000, True)])
def test_host_builder_makes_a_valid_tree(n, clustered):
    info = device.test_bvh_build(random_boxes(n, n, clustered))
    assert info["leaves"] >= (n + 3) // 4 and info["nodes"] >= 1


def test_degenerate_inputs_host():
    b = np.zeros((100, 6), np.float32)                      # identical, zero-sized boxes
    assert device.test_bvh_build(b)["leaves"] >= 25
    b = random_boxes(64, 1); b[:, 2] = 0; b[:, 5] = 0       # flat in z
    device.test_bvh_build(b)


def test_cache_round_trip_and_rejection_host(tmp_path):
    boxes = random_boxes(5000, 3)
    info = device.test_bvh_build(boxes, cache_dir=tmp_path)   # store, load back, compare bit for bit; wrong hash is refused
    files = list(tmp_path.glob("bvh8_*.bin"))
    assert len(files) == 1 and files[0].stat().st_size == 40 + 256 * info["nodes"] + 4 * 5000
    assert not list(tmp_path.glob("*.tmp*"))                   # written under a temporary name, then renamed
    with pytest.raises(device.DeviceError, match="cannot store"):
        device.test_bvh_build(boxes, cache_dir=tmp_path / "does" / "not" / "exist")


@pytest.mark.gpu
@pytest.mark.parametrize("n,clustered", [(5, False), (6, False), (33, False), (1000, False), (100000, False), (50000, True), (2_000_000, False)])
def test_gpu_builder_makes_a_valid_tree(n, clustered, tmp_path):
    boxes = random_boxes(n, n, clustered)
    with device.B200Device() as dev:
        gpu = device.test_bvh_build(boxes, builder=1, device=dev, cache_dir=tmp_path if n <= 100000 else None)
    assert gpu["leaves"] >= (n + 3) // 4
    if n <= 100000:
        host = device.test_bvh_build(boxes)
        # an LBVH is a worse tree than binned SAH, but not wildly: bound the damage (measured 1.1 - 1.6 x)
        assert gpu["sah_x1000"] <= 3 * host["sah_x1000"] + 8000, (gpu, host)


@pytest.mark.gpu
def test_gpu_builder_degenerate_inputs():
    with device.B200Device() as dev:
        b = np.zeros((1000, 6), np.float32)                  # every Morton code equal: the index breaks the ties
        assert device.test_bvh_build(b, builder=1, device=dev)["leaves"] >= 250
        b = random_boxes(4096, 2); b[:, 0] = b[:, 3] = 1.0   # flat in x
        device.test_bvh_build(b, builder=1, device=dev)
        with pytest.raises(device.DeviceError, match="fewer than five"):
            device.test_bvh_build(random_boxes(4, 0), builder=1, device=dev)
