"""CPU tier: the N > 1 host logic of bench.py (slicing, max-over-ranks timing, result gather) with two gloo
ranks -- the solve path itself has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mpc_ros_b200.sharding import rank_slice, reduce_over_ranks, gather_slices


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = rank_slice(total, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64) * 2.0          # stand-in for this rank's solve results
    elapsed = 10.0 + 5.0 * rank                                      # rank 1 is the slow one
    t, (n, it) = reduce_over_ranks(elapsed, [hi - lo, 7.0 * (hi - lo)])
    full = gather_slices(local, total, rank, world)
    dist.barrier()
    q.put((rank, lo, hi, t, n, it, full.tolist()))
    dist.destroy_process_group()


def test_rank_slices_cover_the_batch():
    for total in (0, 1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            prev = 0
            for r in range(world):
                lo, hi = rank_slice(total, r, world)
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == total
            sizes = [rank_slice(total, r, world)[1] - rank_slice(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_gloo():
    world, total = 2, 11
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lo, hi, t, n, it, full in res:
        assert t == 15.0                       # max over ranks
        assert n == total and it == 7.0 * total
        assert full == [2.0 * i for i in range(total)]
    assert res[0][1:3] == (0, 6) and res[1][1:3] == (6, 11)
