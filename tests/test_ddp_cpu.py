"""World-size-2 gloo test of the multi-GPU host logic: gradients are SUMMED across ranks (the loss is a
sum over samples) and every rank ends with identical parameters."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pe_b200 import ddp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    flat = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    red = ddp.BucketedAllReduce(flat, bucket_elems=256, group=dist.group.WORLD, async_op=False)
    # backward-ordered readiness: tail of the arena first
    for lo, hi in [(700, 1000), (300, 700), (0, 300)]:
        red.ready(lo, hi)
    red.wait()
    want = torch.arange(1000, dtype=torch.float32) * sum(r + 1 for r in range(world))
    ok = torch.equal(flat, want)
    lo, hi = ddp.shard_range(10, rank, world)
    out[rank] = (ok, lo, hi, red.launched)
    dist.destroy_process_group()


def test_bucketed_sum_allreduce_world2():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world))
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 5, 5, 10)
    assert out[0][3] >= 3          # 1000 elements in <=256-element buckets, flushed as ranges complete


def test_shard_range_covers_everything():
    for n in (1, 7, 64, 257):
        for world in (1, 2, 4, 8):
            spans = [ddp.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
