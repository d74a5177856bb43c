"""world_size-2 gloo test of the N > 1 host logic: shard ownership tiles the stream ids exactly, and the step time is
the max over ranks.  No GPU, no data-path collective (there is none in the design)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lc3_codec_b200.sharding import max_over_ranks, owner_of, shard_range


def test_shard_range_tiles_exactly():
    for total in (0, 1, 7, 32768, 262144, 65537):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                s, n = shard_range(total, r, world)
                cover.extend(range(s, s + n))
                assert all(owner_of(x, total, world) == r for x in (s, s + n - 1) if n)
            assert cover == list(range(total))
            sizes = [shard_range(total, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, n = shard_range(total, rank, world)
    mine = torch.tensor([s, n], dtype=torch.int64)
    allr = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allr, mine)
    t = max_over_ranks(1.0 + rank, dist)
    q.put((rank, [a.tolist() for a in allr], t))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world = 262144 + 5, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ranges, t in res:
        assert t == 2.0                                   # max over ranks of (1 + rank)
        assert ranges[0][0] == 0 and ranges[0][0] + ranges[0][1] == ranges[1][0]
        assert ranges[1][0] + ranges[1][1] == total
