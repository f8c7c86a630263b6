"""One KernelEstimation head tail (25 -> 25 3x3 at 1088x1920, planar output) for ncu: python tools/prof_head.py [act] [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv as tc
act = None if len(sys.argv) < 2 or sys.argv[1] == "none" else sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with torch.no_grad():
    x = torch.randn((B, 32, 1088, 1920), device="cuda").contiguous(memory_format=torch.channels_last)
    x[:, 25:] = 0
    w = torch.randn((25, 25, 3, 3), device="cuda") / 15
    b = torch.randn(25, device="cuda")
    for _ in range(2):
        y = tc.conv2d(x, w, b, "zeros", act, nchw_out=True)
    torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
