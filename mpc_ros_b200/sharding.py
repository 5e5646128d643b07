"""mpc_ros_b200/sharding.py -- host-side multi-GPU plumbing used by bench.py (one process per GPU).

The solve path shards trivially (independent problems, no exchange step): GPU g owns the
contiguous slice [g*B/G, (g+1)*B/G) of a batch, or -- in the weak-scaling benchmark -- its own batch.  The
only cross-rank operations are the barrier around the timed region, the max over ranks of the device time
and the sum over ranks of the problem counts.  They are backend-agnostic (NCCL on GPUs, gloo in the CPU
tests) so the logic is covered without a GPU.
"""
import torch
import torch.distributed as dist


def rank_slice(total, rank, world):
    """Contiguous slice [lo, hi) of `total` problems owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def reduce_over_ranks(elapsed_ms, counts, device="cpu"):
    """(max over ranks of elapsed_ms, sum over ranks of each entry of counts)."""
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(x) for x in counts], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return float(t.item()), [float(x) for x in c.tolist()]


def gather_slices(local, total, rank, world, device="cpu"):
    """Final host-side gather of per-rank result slices (1-D float64 tensors) into one array on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local.clone()
    sizes = [rank_slice(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, dtype=local.dtype, device=device)
    pad[: local.numel()] = local
    outs = [torch.zeros(mx, dtype=local.dtype, device=device) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(outs, sizes)])
