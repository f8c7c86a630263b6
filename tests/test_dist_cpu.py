"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding without collectives and the
flat-bucket gradient all-reduce of the FusionNet training step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_range_partitions():
    from fvfi.dist import shard_range
    for n in (0, 1, 7, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "fusion-method-for-video-frame-interpolation_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fvfi.dist import FlatGradBucket, broadcast_module, shard_range
    from oracle.nets import FusionNet            # CPU stand-in with FusionNet's parameters: the product module is CUDA-only
    torch.manual_seed(rank)                      # ranks start different; the broadcast makes them rank 0's
    net = FusionNet()
    broadcast_module(net, 0)
    for n, p in net.named_parameters():
        p.requires_grad_(not n.startswith("net."))
    bucket = FlatGradBucket([p for n, p in net.named_parameters() if not n.startswith("net.")])
    g = torch.Generator().manual_seed(1)
    ins = [torch.rand((4, c, 16, 16), generator=g) for c in (3, 3, 3, 6, 3)]
    target = torch.rand((4, 3, 16, 16), generator=g)
    b, e = shard_range(4, rank, world)
    bucket.zero()
    pred = net(*[t[b:e] for t in ins])
    torch.nn.functional.l1_loss(target[b:e], torch.clip(pred, 0, 1)).backward()
    flat = bucket.all_reduce_mean().clone()
    if rank == 0:
        q.put(flat)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_bucket_allreduce_matches_full_batch_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process full-batch reference
    from fvfi.dist import FlatGradBucket
    from oracle.nets import FusionNet
    torch.manual_seed(0)
    net = FusionNet()
    for n, p in net.named_parameters():
        p.requires_grad_(not n.startswith("net."))
    bucket = FlatGradBucket([p for n, p in net.named_parameters() if not n.startswith("net.")])
    g = torch.Generator().manual_seed(1)
    ins = [torch.rand((4, c, 16, 16), generator=g) for c in (3, 3, 3, 6, 3)]
    target = torch.rand((4, 3, 16, 16), generator=g)
    torch.nn.functional.l1_loss(target, torch.clip(net(*ins), 0, 1)).backward()
    assert bucket.flat.numel() == 543331
    assert float((bucket.flat - flat).abs().max()) <= 1e-6
