"""Kernel-time table of one pipeline step (torch.profiler, CUDA activities).  python tools/profile_pipeline.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi import synth as fp   # seeded weights / frames (input generation only)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.backends.cudnn.allow_tf32 = False
pipe = FusionPipeline(1080, 1920, "cuda", phase_plane_chunk=12)
pipe.load_state(fp.seeded_state(0))
r1, r2 = fp.seeded_frames(1, 1080, 1920, 0)
d1, d2 = r1.expand(B, -1, -1, -1).contiguous().cuda(), r2.expand(B, -1, -1, -1).contiguous().cuda()
for _ in range(2):
    pipe(d1, d2)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pipe(d1, d2)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
