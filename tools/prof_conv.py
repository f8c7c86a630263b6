"""One conv layer for ncu: python tools/prof_conv.py Cin Cout K H W [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv
ci, co, k, h, w = [int(v) for v in sys.argv[1:6]]
b = int(sys.argv[6]) if len(sys.argv) > 6 else 2
x = torch.randn((b, ci, h, w), device="cuda").contiguous(memory_format=torch.channels_last)
wt = torch.randn((co, ci, k, k), device="cuda") / (ci * k * k) ** 0.5
bias = torch.randn((co,), device="cuda")
for _ in range(3):
    y = conv.conv2d(x, wt, bias, "zeros", "relu")
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
