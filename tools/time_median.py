"""50x50 exact median on 8 maps of 1080x1920 (the pipeline's call): ms per map.  python tools/time_median.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import filters
x = torch.rand((8, 1080, 1920), device="cuda")
for _ in range(2):
    y = filters.median_filter(x, 50)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    y = filters.median_filter(x, 50)
e1.record()
torch.cuda.synchronize()
print("median 50x50: %.3f ms per 1080p map" % (e0.elapsed_time(e1) / 3 / 8))
