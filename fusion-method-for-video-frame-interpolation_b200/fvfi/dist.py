"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
box, gloo in the CPU tests).

* Inference: frame pairs are independent -> ``shard_range`` splits the batch by rank, NO collective on
  the data path (results are gathered host-side only if the caller wants them in one place).
* Training (FusionNet, config 5): data parallel; ONE flat fp32 bucket holding the gradients of the live
  parameters (543,331 values, 2.17 MB) is all-reduced per step -- latency-bound, so a single message, no
  bucketing/overlap machinery.  The dead ``net.*`` parameters never get gradients and are not in the bucket.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced split of ``n_items`` independent frame pairs: returns (begin, end) of this rank."""
    base, rem = divmod(int(n_items), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_module(module, src=0, group=None):
    """Rank ``src``'s parameters AND buffers (BatchNorm running statistics) to every rank -- what DistributedDataParallel does at
    construction, so replicas start identical whatever each rank loaded or initialised.  One flat message per dtype."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t.numel() > 0]
    for dtype in sorted({t.dtype for t in tensors}, key=str):
        ts = [t for t in tensors if t.dtype == dtype]
        flat = torch.cat([t.detach().reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        with torch.no_grad():
            for t in ts:
                t.copy_(flat[off:off + t.numel()].view_as(t))
                off += t.numel()


def all_reduce_mean_buffers(module, group=None):
    """Average the floating-point buffers (BatchNorm running mean / variance) across ranks after a training step, so the
    statistics every replica would save are the same."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    bufs = [b for b in module.buffers() if b.is_floating_point() and b.numel() > 0]
    if not bufs:
        return
    flat = torch.cat([b.detach().reshape(-1).float() for b in bufs])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    off = 0
    with torch.no_grad():
        for b in bufs:
            b.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()


class FlatGradBucket:
    """Views the ``.grad`` of ``params`` into one contiguous fp32 buffer so a single all-reduce moves them."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)   # grads accumulate straight into the bucket
            off += p.numel()

    def zero(self):
        """Use this instead of ``optimizer.zero_grad()``: ``zero_grad(set_to_none=True)`` would detach ``p.grad`` from the bucket."""
        self.flat.zero_()

    def attached(self):
        """True while every parameter's ``.grad`` still aliases its slice of the flat buffer."""
        off = 0
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                return False
            off += p.numel()
        return True

    def all_reduce_mean(self, group=None):
        assert self.attached(), ("a parameter's .grad no longer points into the flat bucket (optimizer.zero_grad(set_to_none=True)?): "
                                 "the all-reduce would move stale values -- clear gradients with FlatGradBucket.zero()")
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(world)
        return self.flat
