"""FusionNet training step (BASELINE.json configs[4]: 256x256 crops, 8 per GPU, flat-bucket gradient all-reduce over NCCL).
    python tools/bench_train.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py
Prints step time (CUDA events, max over ranks) and the all-reduce share."""
import os, sys, time
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi.trainer import FusionTrainer
from fvfi import synth as fp   # seeded weights / frames (input generation only)

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = W = 256
B = 8
pipe = FusionPipeline(H, W, "cuda")
pipe.load_state(fp.seeded_state(0))
tr = FusionTrainer(pipe, lr=1e-4)
r1, r2 = fp.seeded_frames(B, H, W, rank)
f1, f2 = r1.cuda(), r2.cuda()
target = (0.5 * (f1 + f2)).clamp(0, 1)
for _ in range(3):
    loss = tr.step(f1, f2, target)
torch.cuda.synchronize()
if world > 1:
    dist.barrier(device_ids=[local])
steps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    loss = tr.step(f1, f2, target)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
# all-reduce alone
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a0.record()
for _ in range(steps):
    tr.bucket.all_reduce_mean(tr.group)
a1.record()
torch.cuda.synchronize()
ar = torch.tensor([a0.elapsed_time(a1) / steps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(ar, op=dist.ReduceOp.MAX)
if rank == 0:
    print("train step: %d GPU(s) x %d crops of %dx%d: %.2f ms/step (%.1f crops/s), gradient all-reduce of %d floats %.3f ms, loss %.5f"
          % (world, B, H, W, float(ms), world * B / float(ms) * 1e3, tr.bucket.flat.numel(), float(ar), float(loss)))
if world > 1:
    dist.destroy_process_group()
