"""Micro-benchmark of the tcgen05 3xTF32 convolution on the layer shapes of the three networks (1080p, batch 1-4).
    python tools/bench_conv.py [B]
Prints ms, fp32-equivalent TFLOP/s and the same layer through cuDNN (fp32 and TF32)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fusion-method-for-video-frame-interpolation_b200")]
from fvfi import conv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
LAYERS = [  # name, Cin, Cout, K, H, W, mode
    ("KE conv1b 32->32 full", 32, 32, 3, 1088, 1920, "zeros"),
    ("KE subnet 64->64 half", 64, 64, 3, 544, 960, "zeros"),
    ("KE head 25->25 full", 25, 25, 3, 1088, 1920, "zeros"),
    ("KE conv3 128->128 1/4", 128, 128, 3, 272, 480, "zeros"),
    ("KE conv4 256->256 1/8", 256, 256, 3, 136, 240, "zeros"),
    ("KE conv5 512->512 1/16", 512, 512, 3, 68, 120, "zeros"),
    ("PN 88->64 3x3 refl full", 88, 64, 3, 1080, 1920, "reflect"),
    ("PN 64->64 3x3 refl full", 64, 64, 3, 1080, 1920, "reflect"),
    ("PN 64->8 1x1 full", 64, 8, 1, 1080, 1920, "zeros"),
    ("FN 18->32 5x5 refl full", 18, 32, 5, 1080, 1920, "reflect"),
    ("FN 32->64 5x5 refl half", 32, 64, 5, 540, 960, "reflect"),
]


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, ci, co, k, h, w, mode in LAYERS:
    b = 1 if h * w > 1500000 and max(ci, co) >= 64 else B
    x = torch.randn((b, ci, h, w), device="cuda").contiguous(memory_format=torch.channels_last)
    wt = torch.randn((co, ci, k, k), device="cuda") / (ci * k * k) ** 0.5
    bias = torch.randn((co,), device="cuda")
    flops = 2.0 * b * h * w * ci * co * k * k
    conv.precision = conv.PRECISIONS["tf32x3"]
    t_t3 = timeit(lambda: conv.conv2d(x, wt, bias, mode, "relu"))
    conv.precision = conv.PRECISIONS["f16x3"]
    t_tc = timeit(lambda: conv.conv2d(x, wt, bias, mode, "relu"))
    p = k // 2
    def cudnn():
        xx = F.pad(x, (p, p, p, p), mode="reflect") if (mode == "reflect" and p) else x
        return F.relu(F.conv2d(xx, wt, bias, padding=0 if (mode == "reflect" and p) else p))
    torch.backends.cudnn.allow_tf32 = False
    t_f32 = timeit(cudnn)
    torch.backends.cudnn.allow_tf32 = True
    t_tf32 = timeit(cudnn)
    print("%-26s B=%d  tcgen05 3xFP16 %7.3f ms %6.1f TF/s | 3xTF32 %7.3f ms %6.1f | cuDNN fp32 %7.3f ms %6.1f | cuDNN tf32 %7.3f ms %6.1f" %
          (name, b, t_tc, flops / t_tc / 1e9, t_t3, flops / t_t3 / 1e9, t_f32, flops / t_f32 / 1e9, t_tf32, flops / t_tf32 / 1e9))
