"""World-size-2 gloo test of the clip sharding / result gather used for multi-GPU runs."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openvis_b200.sharding import gather_clip_dict, gather_clip_results, shard_range


def test_shard_range_covers_everything():
    for n in (1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            got = [i for r in range(world) for i in shard_range(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_range(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = shard_range(n_items, rank, world)
        # "result" of clip c: a [3, 2] tensor filled with c (stands for per-clip class scores)
        local = torch.stack([torch.full((3, 2), float(c)) for c in mine]) if len(mine) else torch.zeros(0, 3, 2)
        full = gather_clip_results(local, n_items)
        ok = full.shape == (n_items, 3, 2) and all(bool((full[c] == c).all()) for c in range(n_items))
        # heterogeneous per-clip results of the online models: scores, query-matching indices, packed masks
        loc = {"scores": local, "indices": torch.stack([torch.full((4, 5), c, dtype=torch.int16) for c in mine]) if len(mine)
               else torch.zeros(0, 4, 5, dtype=torch.int16),
               "bits": torch.stack([torch.full((2, 3), -c - 1, dtype=torch.int32) for c in mine]) if len(mine)
               else torch.zeros(0, 2, 3, dtype=torch.int32)}
        d = gather_clip_dict(loc, n_items)
        ok = ok and d["indices"].dtype == torch.int16 and d["bits"].dtype == torch.int32
        ok = ok and all(bool((d["indices"][c] == c).all()) and bool((d["bits"][c] == -c - 1).all()) for c in range(n_items))
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    for n_items in (5, 4):
        ret = ctx.Manager().dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret[0] and ret[1]
        port += 1
