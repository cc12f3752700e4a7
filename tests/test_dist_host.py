"""Host-side logic of the strip partition on CPU: strip bounds, pixel ranges, and the rank-ordered
blob exchange over torch.distributed (gloo, world_size 2).  The CUDA data path is covered by
tests/dist_strip_check.py on a multi-GPU box (test_gpu_dist.py launches it when >= 2 GPUs exist)."""
import os
import socket

import numpy as np
import pytest

from oracle import srps_oracle as o
from srmeetsps_cuda_b200.dist import exchange_blobs, local_ranges, strip_bounds


def test_strip_bounds_are_aligned_and_cover():
    for w, world in [(4096, 8), (4096, 2), (48, 2), (1920, 7), (128, 3)]:
        b = strip_bounds(w, world)
        assert b[0][0] == 0 and b[-1][1] == w and len(b) == world
        for (a0, a1), (b0, b1) in zip(b[:-1], b[1:]):
            assert a1 == b0
        assert all(x % 4 == 0 and y % 4 == 0 and y > x for x, y in b)
        sizes = [y - x for x, y in b]
        assert max(sizes) - min(sizes) <= 4
    with pytest.raises(ValueError):
        strip_bounds(8, 3)


def test_local_ranges_match_masked_order():
    sc = o.synth_scene(40, 48, 2, 3, seed=5, mask_kind="random95")
    ops = sc["ops"]
    h = 40
    for (j0, j1) in strip_bounds(48, 3):
        p0, p1, q0, q1 = local_ranges(sc["mask"], 2, j0, j1)
        cols = ops["imask"] // h
        assert p0 == int((cols < j0).sum()) and p1 == int((cols < j1).sum())
        lcols = ops["imasks"] // (h // 2)
        assert q0 == int((lcols < j0 // 2).sum()) and q1 == int((lcols < j1 // 2).sum())


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blobs = exchange_blobs(bytes([rank]) * 16)
    q.put((rank, blobs))
    dist.destroy_process_group()


def test_blob_exchange_is_rank_ordered_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    for rank, blobs in res:
        assert blobs == [bytes([0]) * 16, bytes([1]) * 16]


def test_synthetic_strips_are_slices_of_the_whole_scene():
    """bench.py's scene generator: a strip of a scene holds exactly the values the whole scene has on those columns
    (counter-based noise, apron for the z initialisation), so every GPU count solves the same scene."""
    from srmeetsps_cuda_b200.synth import hash_normal, synth_scene_torch
    import torch
    x = hash_normal(torch.arange(0, 400000, dtype=torch.int64), 2000, 3)
    assert abs(float(x.mean())) < 5e-3 and abs(float(x.std()) - 1.0) < 5e-3
    h, w, sf, n = 64, 96, 4, 5
    full = synth_scene_torch(h, w, sf, n, 2000, device="cpu", pin=False)
    for world in (2, 3, 8):
        parts = [synth_scene_torch(h, w, sf, n, 2000, device="cpu", pin=False, j0=j0, j1=j1) for j0, j1 in strip_bounds(w, world)]
        assert np.array_equal(np.concatenate([p["z"] for p in parts]), full["z"])
        assert np.array_equal(np.concatenate([p["z0s"] for p in parts]), full["z0s"])
        assert np.array_equal(np.concatenate([p["I"] for p in parts], axis=2), full["I"])
