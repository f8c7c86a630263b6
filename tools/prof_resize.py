"""The two dominant bilinear-resize shapes of the pipeline, for ncu / timing: python tools/prof_resize.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv as tc
def t(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# (a) KernelEstimation head tail: [8,32,544,960] -> x2, align_corners=True
x = torch.rand((8, 32, 544, 960), device="cuda").contiguous(memory_format=torch.channels_last)
ms = t(lambda: tc.resize_bilinear(x, (1088, 1920), True))
gb = (x.numel() + 8 * 32 * 1088 * 1920) * 4 / 1e9
print("head tail x2 (32 ch, align_corners): %.3f ms, %.2f GB -> %.0f GB/s" % (ms, gb, gb / ms * 1e3))
# (b) PhaseNet level 0: feature [12,64,764,1358] -> [1080,1920] into channels 0..63 of an 88-channel concat
f = torch.rand((12, 64, 764, 1358), device="cuda").contiguous(memory_format=torch.channels_last)
cat = torch.empty((12, 88, 1080, 1920), device="cuda").contiguous(memory_format=torch.channels_last)
ms = t(lambda: tc.resize_bilinear(f, (1080, 1920), False, out=cat, out_channel_offset=0), 3)
gb = (f.numel() + 12 * 64 * 1080 * 1920) * 4 / 1e9
print("PhaseNet level 0 (64 of 88 ch, ratio sqrt2): %.3f ms, %.2f GB -> %.0f GB/s" % (ms, gb, gb / ms * 1e3))
# (c) FusionNet decoder: [8,64,540,960] relu + x2 + skip
y = torch.rand((8, 64, 540, 960), device="cuda").contiguous(memory_format=torch.channels_last)
s = torch.rand((8, 64, 1080, 1920), device="cuda").contiguous(memory_format=torch.channels_last)
ms = t(lambda: tc.resize_bilinear(y, (1080, 1920), False, relu_input=True, add=s), 3)
gb = (y.numel() + 2 * s.numel()) * 4 / 1e9
print("FusionNet decoder step (64 ch, relu in, skip add): %.3f ms, %.2f GB -> %.0f GB/s" % (ms, gb, gb / ms * 1e3))
