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


class OverlappedGather:
    """The path's one exchange, taken off the critical path: every step's [B,H,W] disparity maps are all-gathered on a side
    stream while the next step's towers already run, instead of a rendez-vous of all ranks at the end of every step
    (with a blocking gather the slowest GPU of the step sets everybody's pace; round 1 measured 0.956 scaling at N = 8).

    `depth` output buffers rotate: the engine writes step i into buffer i % depth, which is free again as soon as the
    gather of step i - depth has finished (enforced with an event, never by the host).

        g = OverlappedGather((B, H, W), torch.float32, device)
        for ...:
            out = g.next_buffer()          # where this step's disparities go
            engine(left, right, out=out)
            everyone = g.submit()          # [world*B,H,W]; valid once g.wait(everyone) / g.flush() was called
        g.flush()

    On a CPU process group (gloo, used by the tests) there are no streams: submit() gathers synchronously."""

    def __init__(self, shape, dtype, device, depth=2):
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.cuda = torch.device(device).type == "cuda"
        self.depth = depth
        self.local = [torch.empty(tuple(shape), dtype=dtype, device=device) for _ in range(depth)]
        self.out = [torch.empty((self.world * shape[0],) + tuple(shape[1:]), dtype=dtype, device=device) for _ in range(depth)]
        self.done = [None] * depth
        self.side = torch.cuda.Stream(device=device) if self.cuda else None
        self.step = 0

    def next_buffer(self):
        k = self.step % self.depth
        if self.cuda and self.done[k] is not None:
            torch.cuda.current_stream().wait_event(self.done[k])      # its previous gather must have read it
        return self.local[k]

    def submit(self):
        k = self.step % self.depth
        self.step += 1
        if self.world == 1:
            self.out[k] = self.local[k]
            return self.out[k]
        if not self.cuda:
            gather_disparities(self.local[k], out=self.out[k])
            return self.out[k]
        ready = torch.cuda.Event()
        ready.record()                                                # this step's soft-argmin has been enqueued
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_gather_into_tensor(self.out[k], self.local[k])
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.done[k] = ev
        return self.out[k]

    def flush(self):
        """Make the current stream wait for every gather issued so far."""
        if self.cuda:
            for ev in self.done:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
