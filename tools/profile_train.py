"""Kernel-time table of one FusionNet training step (BASELINE.json configs[4], 8 crops of 256x256 on one GPU)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi.trainer import FusionTrainer
from fvfi import synth as fp
H = W = 256
pipe = FusionPipeline(H, W, "cuda")
pipe.load_state(fp.seeded_state(0))
tr = FusionTrainer(pipe, lr=1e-4, graph_frozen=len(sys.argv) > 1)
r1, r2 = fp.seeded_frames(8, H, W, 0)
f1, f2 = r1.cuda(), r2.cuda()
target = (0.5 * (f1 + f2)).clamp(0, 1)
for _ in range(4):
    tr.step(f1, f2, target)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(f1, f2, target)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=80))
