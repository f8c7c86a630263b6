"""Per-call-site device time of the host-level operators of one pipeline step, grouped by (operator, shapes):
python tools/prof_calls.py [B] [operator substring ...]   (CUDA events around every call of fvfi.conv.resize_bilinear / avg_pool2 /
conv2d, fvfi.filters.*, the pyramid calls ...)"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi import synth as fp, conv as tc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
want = sys.argv[2:] or ["resize_bilinear", "avg_pool2", "max_pool2"]
pipe = FusionPipeline(1080, 1920, "cuda", phase_plane_chunk=12)
pipe.load_state(fp.seeded_state(0))
r1, r2 = fp.seeded_frames(1, 1080, 1920, 0)
d1, d2 = r1.expand(B, -1, -1, -1).contiguous().cuda(), r2.expand(B, -1, -1, -1).contiguous().cuda()
for _ in range(2):
    pipe(d1, d2)
torch.cuda.synchronize()
records = []


def wrap(mod, name):
    fn = getattr(mod, name)

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        shapes = tuple(tuple(t.shape) for t in a if torch.is_tensor(t))
        extra = tuple((kk, tuple(v.shape) if torch.is_tensor(v) else v) for kk, v in sorted(k.items()) if kk != "out") + tuple(
            x for x in a if isinstance(x, (int, bool, tuple, str)))
        records.append((name, shapes, extra, e0, e1))
        return out
    setattr(mod, name, inner)


for n in want:
    if hasattr(tc, n):
        wrap(tc, n)
pipe(d1, d2)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for name, shapes, extra, e0, e1 in records:
    k = (name, shapes, extra)
    v = agg.setdefault(k, [0, 0.0])
    v[0] += 1; v[1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print("B = %d: %d calls, %.2f ms" % (B, len(records), tot))
for (name, shapes, extra), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%7.3f ms  x%-3d %s %s %s" % (ms, cnt, name, shapes, extra))
