"""Data-parallel plumbing: frame tuples are independent, so the batch is split contiguously by
rank (the reference's nn.DataParallel scatter, eval.py:206-207) and the only collectives are one
all-gather of the outputs and one all-reduce of scalar metrics.
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world_size):
    """Contiguous [start, stop) of `total` items for `rank`; the first `total % world_size` ranks get one more."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world_size %d" % (rank, world_size))
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_outputs(local, total, group=None):
    """All-gathers per-rank output blocks [n_local, ...] (sharded with shard_range) into [total, ...]."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(b - a for a, b in sizes)
    pad = local
    if local.shape[0] < nmax:
        pad = torch.cat([local, local.new_zeros((nmax - local.shape[0],) + tuple(local.shape[1:]))], 0)
    out = local.new_empty((world * nmax,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(b - a == nmax for a, b in sizes):
        return out
    return torch.cat([out[r * nmax: r * nmax + (b - a)] for r, (a, b) in enumerate(sizes)], 0)


def reduce_metrics(values, op="sum", group=None):
    """All-reduces a small 1-D tensor of scalars (frame counts, error sums, elapsed time)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return values
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    out = values.clone()
    dist.all_reduce(out, op=ops[op], group=group)
    return out


class OutputGatherer:
    """ONE all-gather per step of a rank's packed output block (depth, normal, prob of its B frames: 5 fp32 maps per
    frame, 20*H*W*B bytes), issued on a side stream so that it overlaps the next step's compute (SURVEY.md section 8e;
    the reference's equivalent is nn.DataParallel's gather, eval.py:206-207).

    submit(packed) is called on the compute stream right after a step: it snapshots the block into a send slot (one
    small device copy), and the side stream all-gathers that slot into the matching receive slot [world, numel].  Two
    slots: the step that reuses a slot first waits for the gather that last read it.  result(slot) makes the calling
    stream wait for that gather and returns the per-rank blocks.  With one process (or no process group) everything is
    a no-op on the input.  Works on CPU tensors too (gloo tests): then there are no streams and submit() blocks."""

    def __init__(self, numel, device, dtype=torch.float32, slots=2, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.slots = slots
        self.i = 0
        if self.world > 1:
            self.send = [torch.empty((numel,), dtype=dtype, device=self.device) for _ in range(slots)]
            self.recv = [torch.empty((self.world * numel,), dtype=dtype, device=self.device) for _ in range(slots)]
            self.numel = numel
            self.done = [None] * slots
            self.side = torch.cuda.Stream(self.device) if self.cuda else None
        self.last = None

    def submit(self, packed):
        """-> slot index (pass it to result())."""
        if self.world == 1:
            self.last = packed
            return 0
        slot = self.i % self.slots
        self.i += 1
        flat = packed.reshape(-1)
        if not self.cuda:
            self.send[slot].copy_(flat)
            dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group)
            return slot
        cur = torch.cuda.current_stream(self.device)
        if self.done[slot] is not None:
            cur.wait_event(self.done[slot])      # the gather that last read this send slot
        self.send[slot].copy_(flat, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ready)
            dist.all_gather_into_tensor(self.recv[slot], self.send[slot], group=self.group)
            self.done[slot] = torch.cuda.Event()
            self.done[slot].record(self.side)
        return slot

    def result(self, slot):
        """[world, numel] blocks of the step submitted into `slot`; the calling stream waits for its gather."""
        if self.world == 1:
            return self.last.reshape(1, -1)
        if self.cuda and self.done[slot] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.done[slot])
        return self.recv[slot].view(self.world, self.numel)

    def join(self):
        """The calling stream waits for every gather issued so far."""
        if self.world > 1 and self.cuda:
            torch.cuda.current_stream(self.device).wait_stream(self.side)


def unpack_outputs(blocks, B, H, W):
    """[world, 5*B*H*W] packed blocks (CNMPipeline's `packed` buffer of every rank) -> depth [world*B,H,W],
    normal [world*B,3,H,W], prob [world*B,1,H,W]."""
    world = blocks.shape[0]
    n = B * H * W
    depth = blocks[:, :n].reshape(world * B, H, W)
    normal = blocks[:, n:4 * n].reshape(world * B, 3, H, W)
    prob = blocks[:, 4 * n:5 * n].reshape(world * B, 1, H, W)
    return depth, normal, prob
