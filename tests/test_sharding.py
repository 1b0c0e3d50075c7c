"""N>1 host logic on CPU: world_size-2 gloo processes shard the frame list and gather timings."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fovgs import shard


def test_round_robin_covers_every_frame_once():
    for n, w in [(270, 1), (270, 2), (270, 4), (270, 8), (7, 8), (0, 2)]:
        seen = sorted(f for r in range(w) for f in shard.frames_for_rank(n, r, w))
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard.frames_for_rank(10, 2, 2)


def test_frame_assignment_is_the_9_gaze_protocol():
    pairs = [p for r in range(8) for p in shard.frame_assignment(30, 9, r, 8)]
    assert sorted(pairs) == sorted((c, g) for g in range(9) for c in range(30))


def test_aggregate_uses_slowest_rank():
    t = torch.tensor([[1000.0, 100.0], [2000.0, 100.0]], dtype=torch.float64)
    assert shard.aggregate_fps(t) == pytest.approx(200 / 2.0)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    frames = shard.frames_for_rank(31, rank, world)
    ms = 10.0 * len(frames) * (1 + rank)          # rank 1 is slower
    table = shard.gather_timings(ms, len(frames))
    imgs = shard.gather_images(torch.full((3, 4, 5), float(rank + 1)))
    ok = (rank != 0 and imgs is None) or (rank == 0 and imgs.shape == (world, 3, 4, 5)
                                          and all(bool((imgs[r] == r + 1).all()) for r in range(world)))
    q.put((rank, frames, table.tolist(), shard.aggregate_fps(table), ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get() for _ in range(2)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == list(range(0, 31, 2)) and res[1][1] == list(range(1, 31, 2))
    assert res[0][2] == res[1][2]                 # both ranks hold the same table
    assert res[0][3] == pytest.approx(31 / (10.0 * 15 * 2 * 1e-3))
    assert res[0][4] and res[1][4]                # rank 0 holds both ranks' images, rank 1 holds none
