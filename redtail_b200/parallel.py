"""Data-parallel plumbing of the stereo path: stereo pairs are independent, so a batch is split into contiguous
blocks over the ranks (one process per GPU) and the only exchange is one all-gather of the disparity maps
(NCCL over NVLink on GPUs; gloo in the CPU tests).  No other collective exists on this path."""
import torch
import torch.distributed as dist


def shard_range(total_pairs, world, rank):
    """Contiguous block [begin, end) of `total_pairs` owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(total_pairs, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_disparities(local, out=None):
    """local: [B,H,W] on every rank (same B)  ->  [world*B,H,W] on every rank, rank-major order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, local.contiguous())
    else:
        dist.all_gather(list(out.chunk(world, dim=0)), local.contiguous())
    return out
