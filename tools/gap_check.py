"""GPU-busy fraction of one pipeline step: event-timed wall time of a step against the sum of its kernel durations
(torch.profiler).  python tools/gap_check.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi import synth as fp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
pipe = FusionPipeline(1080, 1920, "cuda", phase_plane_chunk=12)
pipe.load_state(fp.seeded_state(0))
r1, r2 = fp.seeded_frames(1, 1080, 1920, 0)
d1, d2 = r1.expand(B, -1, -1, -1).contiguous().cuda(), r2.expand(B, -1, -1, -1).contiguous().cuda()
for _ in range(2):
    pipe(d1, d2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    pipe(d1, d2)
e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / 3
pipe.timing = []
pipe(d1, d2); torch.cuda.synchronize()
stages = {}
for (_, a), (name, b) in zip(pipe.timing, pipe.timing[1:]):
    if name != 'start':
        stages[name] = stages.get(name, 0.0) + a.elapsed_time(b)
pipe.timing = None
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    pipe(d1, d2)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
busy = sum(e.device_time for e in evs) / 1e3
print("B = %d: step %.1f ms (events), kernels %.1f ms (%d launches) -> GPU busy %.1f %%" % (B, wall, busy, len(evs), 100 * busy / wall))
# idle time attributed to the stage windows: sort kernels by start, accumulate gaps > 5 us
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
gaps = []
for a, b in zip(evs, evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 5:
        gaps.append((g, a.name[:60], b.name[:60], (a.time_range.end - t0) / 1e3))
print("idle in gaps > 5 us: %.1f ms in %d gaps" % (sum(g[0] for g in gaps) / 1e3, len(gaps)))
for g in sorted(gaps, reverse=True)[:25]:
    print("  %7.1f us at %7.1f ms  after %-60s before %s" % (g[0], g[3], g[1], g[2]))
print({k: round(v, 1) for k, v in stages.items()})
