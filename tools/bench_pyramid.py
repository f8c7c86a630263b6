"""Pyramid decompose / reconstruct timing at 1080p.  python tools/bench_pyramid.py [N planes] [reps] [H W height]"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi.pyramid import Pyramid
N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, W, height = (int(v) for v in sys.argv[3:6]) if len(sys.argv) > 5 else (1080, 1920, 17)
pyr = Pyramid(height, 4, np.sqrt(2), torch.device("cuda"))
x = torch.rand((N, H, W), device="cuda")
def t(fn, reps=reps):
    fn(); fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
vals = pyr.filter(x)
t(lambda: pyr.filter(x))
td = t(lambda: pyr.filter(x))
tdn = t(lambda: pyr.filter(x, want_high=False))
tr = t(lambda: pyr.inv_filter(vals))
trs = t(lambda: pyr.inv_filter_sparse(vals, use_high=False))
bytes_plane = 72 * H * W
print("decompose   %.3f ms total, %.3f ms/plane, %.1f GB/s algorithmic" % (td, td / N, N * bytes_plane / td / 1e6))
print("decompose (no high) %.3f ms total, %.3f ms/plane" % (tdn, tdn / N))
print("reconstruct %.3f ms total, %.3f ms/plane, %.1f GB/s algorithmic" % (tr, tr / N, N * bytes_plane / tr / 1e6))
print("reconstruct (no high) %.3f ms total, %.3f ms/plane" % (trs, trs / N))
