"""world_size-2 gloo tests of the sharding / broadcast / gather plumbing used for N > 1 GPUs."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from resshift_b200.parallel import shard_range, broadcast_state_dict, gather_shards


def test_shard_range_matches_reference_slicing():
    # reference sampler.py:273-277: micro = ceil(bs / G); rank r keeps [r*micro, (r+1)*micro)
    assert [shard_range(16, 8, r) for r in range(8)] == [(2 * r, 2 * r + 2) for r in range(8)]
    assert [shard_range(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert shard_range(1, 4, 3) == (1, 1)            # trailing ranks may be empty
    for bs in range(1, 20):
        for g in (1, 2, 3, 4, 8):
            covered = [i for r in range(g) for i in range(*shard_range(bs, g, r))]
            assert covered == list(range(bs))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(rank)                                   # different weights per rank before broadcast
        sd = {"a.weight": torch.randn(4, 3), "b.bias": torch.randn(5), "idx": torch.arange(3)}
        broadcast_state_dict(sd, src=0)
        torch.manual_seed(0)
        ref = {"a.weight": torch.randn(4, 3), "b.bias": torch.randn(5)}
        ok = all(torch.equal(sd[k], ref[k]) for k in ref)
        batch = 5
        s, e = shard_range(batch, world, rank)
        local = torch.arange(s, e, dtype=torch.float32)[:, None].repeat(1, 3)   # "image" i is filled with i
        full = gather_shards(local, batch)
        ok = ok and torch.equal(full[:, 0], torch.arange(batch, dtype=torch.float32)) and full.shape == (batch, 3)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
