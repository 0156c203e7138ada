"""CPU, world_size 2, gloo: the N>1 host logic of the path (batch sharding + the one disparity all-gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from redtail_b200.parallel import OverlappedGather, shard_range, gather_disparities


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(total, world, rank)
    # fake "disparity maps": pair i is filled with the value i
    local = torch.stack([torch.full((5, 7), float(i)) for i in range(b, e)])
    full = gather_disparities(local)
    # the rotating-buffer variant bench.py uses (synchronous on a CPU group): three steps through two buffers
    g = OverlappedGather(tuple(local.shape), local.dtype, "cpu", depth=2)
    rot = []
    for step in range(3):
        g.next_buffer().copy_(local + 100.0 * step)
        rot.append(g.submit()[:, 0, 0].tolist())
    g.flush()
    q.put((rank, (b, e), full[:, 0, 0].tolist(), rot))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_batch():
    for total in (1, 2, 7, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_gloo():
    world, total = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, span, order, rot in res:
        assert order == [float(i) for i in range(total)], (rank, order)     # every rank sees all pairs in batch order
        for step, got in enumerate(rot):
            assert got == [float(i) + 100.0 * step for i in range(total)], (rank, step, got)


def test_single_process_is_identity():
    x = torch.randn(2, 3, 4)
    assert gather_disparities(x) is x
    g = OverlappedGather((2, 3, 4), torch.float32, "cpu")
    buf = g.next_buffer()
    buf.copy_(x)
    assert torch.equal(g.submit(), x)
