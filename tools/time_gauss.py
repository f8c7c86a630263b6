"""scipy-equivalent Gaussian (sigma = 5) on 8 maps of 1080x1920 (the pipeline's call): ms per map.  python tools/time_gauss.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import filters
x = torch.rand((8, 1080, 1920), device="cuda")
for sigma in (5.0, 6.0):
    for _ in range(2):
        y = filters.gaussian_filter(x, sigma)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = filters.gaussian_filter(x, sigma)
    e1.record()
    torch.cuda.synchronize()
    print("gaussian sigma %.1f (%s kernels): %.3f ms per 1080p map" % (sigma, "register-weight" if sigma == 5.0 else "generic", e0.elapsed_time(e1) / 5 / 8))
