"""Upsample -> Conv2d fused (fvfi_conv2d_nhwc_upsampled) against resize kernel + convolution on the pipeline's shapes:
python tools/bench_upconv.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv


def t(fn, reps=5):
    for _ in range(2): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name, cin, cs, cout, hs, ws, act, nchw in [("head tail 25->25 (planar out)", 25, 32, 25, 544, 960, None, True),
                                               ("head tail 25->25 softmax", 25, 32, 25, 544, 960, "softmax", True),
                                               ("moduleUpsample2 64->64", 64, 64, 64, 272, 480, "relu", False),
                                               ("moduleUpsample3 128->128", 128, 128, 128, 136, 240, "relu", False),
                                               ("moduleUpsample4 256->256", 256, 256, 256, 68, 120, "relu", False),
                                               ("moduleUpsample5 512->512", 512, 512, 512, 34, 60, "relu", False)]:
    x = torch.randn((B, cs, hs, ws), device="cuda").contiguous(memory_format=torch.channels_last)
    w = torch.randn((cout, cin, 3, 3), device="cuda") / (3 * cin ** 0.5)
    b = torch.randn((cout,), device="cuda")
    size = (2 * hs, 2 * ws)
    t_res = t(lambda: conv.resize_bilinear(x, size, True))
    up = conv.resize_bilinear(x, size, True)
    t_conv = t(lambda: conv.conv2d(up, w, b, "zeros", act, nchw_out=nchw))
    t_fused = t(lambda: conv.conv2d(x, w, b, "zeros", act, nchw_out=nchw, upsample=(size, True)))
    same = torch.equal(conv.conv2d(up, w, b, "zeros", act, nchw_out=nchw), conv.conv2d(x, w, b, "zeros", act, nchw_out=nchw, upsample=(size, True)))
    print("%-30s resize %.3f + conv %.3f = %.3f ms   fused %.3f ms   (identical: %s)" % (name, t_res, t_conv, t_res + t_conv, t_fused, same))
