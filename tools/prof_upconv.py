"""One fused head tail (Upsample x2 -> Conv2d(25, 25, 3), planar output) for ncu: python tools/prof_upconv.py [B] [fused 0/1]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fused = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = torch.randn((B, 32, 544, 960), device="cuda").contiguous(memory_format=torch.channels_last)
w = torch.randn((25, 25, 3, 3), device="cuda") / 15
b = torch.randn((25,), device="cuda")
for _ in range(3):
    if fused:
        y = conv.conv2d(x, w, b, "zeros", None, nchw_out=True, upsample=((1088, 1920), True))
    else:
        y = conv.conv2d(conv.resize_bilinear(x, (1088, 1920), True), w, b, "zeros", None, nchw_out=True)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
