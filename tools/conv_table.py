"""Per-shape table of the tcgen05 convolutions of one pipeline step (CUDA events per launch): python tools/conv_table.py [B]"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pipeline import FusionPipeline
from fvfi import conv as tc
from fvfi import synth as fp   # seeded weights / frames (input generation only)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
pipe = FusionPipeline(1080, 1920, "cuda", phase_plane_chunk=6)
pipe.load_state(fp.seeded_state(0))
r1, r2 = fp.seeded_frames(1, 1080, 1920, 0)
d1, d2 = r1.expand(B, -1, -1, -1).contiguous().cuda(), r2.expand(B, -1, -1, -1).contiguous().cuda()
for _ in range(2):
    pipe(d1, d2)
torch.cuda.synchronize()
tc.timing = []
pipe(d1, d2)
torch.cuda.synchronize()
rec, tc.timing = tc.timing, None
agg = collections.OrderedDict()
for fl, e0, e1, parts, shape, up in rec:
    a = agg.setdefault(shape + (('up',) if up else ()), [0, 0.0, 0.0])
    a[0] += 1; a[1] += fl; a[2] += e0.elapsed_time(e1)
tot = sum(a[2] for a in agg.values())
print("total conv %.2f ms, %.1f TFLOP/s" % (tot, sum(a[1] for a in agg.values()) / tot / 1e9))
for shape, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:40]:
    print("%-44s x%-3d %8.3f ms %5.1f%%  %6.1f TF/s" % (str(shape), n, ms, 100 * ms / tot, fl / ms / 1e9))
