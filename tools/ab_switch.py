"""Same-process A/B of a boolean switch of fvfi.conv on one pipeline call (1080p, B frame pairs), alternating:
    python tools/ab_switch.py fuse_upsample [B] [rounds] [value ...]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi import synth as fp, conv as tc
name = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pipe = FusionPipeline(1080, 1920, "cuda", phase_plane_chunk=24)
pipe.max_batch = 8
pipe.load_state(fp.seeded_state(0))
r1, r2 = fp.seeded_frames(1, 1080, 1920, 0)
d1, d2 = r1.expand(B, -1, -1, -1).contiguous().cuda(), r2.expand(B, -1, -1, -1).contiguous().cuda()


def run(flag, reps=3):
    setattr(tc, name, flag)
    pipe(d1, d2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = pipe(d1, d2)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


for _ in range(2):
    run(True)
vals = [True, False] if len(sys.argv) <= 4 else [eval(v) for v in sys.argv[4:]]
res = {v: [] for v in vals}
outs = {}
for _ in range(rounds):
    for flag in vals:
        ms, outs[flag] = run(flag)
        res[flag].append(ms)
for flag in vals:
    print("%s = %s: %s  mean %.2f ms" % (name, flag, " ".join("%.2f" % v for v in res[flag]), sum(res[flag]) / len(res[flag])))
print("outputs identical:", all(bool(torch.equal(outs[vals[0]], outs[v])) for v in vals[1:]))
