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
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size(group)
            if world > 1:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(world)
        return self.flat
